"""The radix-refinement nearest_probability_distribution (csrc/npd.cu) restated in numpy
(tests/npd_emulation.py) against the oracle and the reference's golden cases: pins the LOGIC of the
kernels - key ranges, bins, integer quantisation, tails, the multi-rank form - without a GPU."""
import numpy as np
import pytest

import npd_emulation as ne
from conftest import load_golden
from oracle import dense as od


def _check(v, acc=0.0, shards=1, tol=1e-15):
    want = od.nearest_probability_distribution(np.asarray(v, float), acc)
    got11, st11 = ne.nearest_probability_distribution(v, acc, 1, bin_bits=11)     # npd_cluster_kernel's geometry
    assert np.abs(got11 - want).max() < tol and st11.passes <= st11.levels + 1
    got, st = ne.nearest_probability_distribution(v, acc, shards)
    assert np.abs(got - want).max() < tol
    assert st.status == st11.status and (st.status != ne.SOLVED or st.num == st11.num)   # the same partition
    return st


def test_golden_cases():
    for c in load_golden("knit_cases.json")["npd"]:
        dense = np.zeros(1 << c["nbits"])
        for k, v in c["raw"]:
            dense[int(k)] = v
        want = np.zeros_like(dense)
        for k, v in c["out"]:
            want[int(k)] = v
        for shards, bits in ((1, 13), (3, 13), (1, 11)):
            got, st = ne.nearest_probability_distribution(dense, c["acc"], shards, bin_bits=bits)
            assert np.abs(got - want).max() < 1e-13
            assert st.status in (ne.IDENTITY, ne.SOLVED)


@pytest.mark.parametrize("kind", range(5))
def test_random_families(kind):
    rng = np.random.default_rng(10 + kind)
    for trial in range(12):
        n = int(rng.integers(2, 200))
        if kind == 0:                                   # large negative entries
            v = rng.normal(0, 1, n)
            v[0] += abs(v.sum()) + 1
        elif kind == 1:                                 # one peak + rounding noise (bv-like knit results)
            v = np.zeros(n)
            v[rng.integers(n)] = 1.0
            v += rng.normal(0, 1e-17, n)
        elif kind == 2:
            v = rng.random(n)
            v /= v.sum()
            v[:n // 2] -= 1e-3
            if v.sum() < 0:
                v[-1] += 1
        elif kind == 3:                                 # exact ties at and around the threshold
            v = np.full(n, -1e-17)
            v[-1] = 1.0
            v[n // 2] = 3e-17
        else:
            v = rng.choice([-2e-17, -1e-17, 1e-17, 5e-17, 0.0, 0.25], n)
            v[0] = 1.0
        st = _check(v, shards=1 + trial % 3, tol=1e-12 if kind in (0, 2) else 1e-15)
        assert st.passes <= ne.LEVELS + 1


def test_edge_cases():
    assert _check(np.array([0.5, 0.5])).status == ne.IDENTITY
    assert _check(np.array([-0.1, 0.6, 0.5])).status == ne.SOLVED
    _check(np.array([-0.2, 0.05, 1.15]))                # a small positive entry is dropped too
    _check(np.array([-1e-300, 1e-300, 1.0, 5e-301]))    # extreme dynamic range (subnormal bin widths)
    _check(np.array([-0.5, 0.2, 0.2, 0.2, 0.2, 0.2]))
    _, st = ne.nearest_probability_distribution(np.array([-1.0, 0.5]))
    assert st.status == ne.NEGATIVE_TOTAL
    rng = np.random.default_rng(0)
    v = rng.normal(0, 3e-5, 300)
    v[0] = 1.0
    _check(v, acc=1e-5, tol=1e-13)                      # pruned entries neither count nor receive the shift
