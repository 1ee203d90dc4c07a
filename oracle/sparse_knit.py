"""Sparse quasi-distribution algebra and the knit driver (TEST INFRASTRUCTURE).

Restates, as plain functions over ``dict[int, float]`` with an explicit
``acc`` threshold instead of the reference's module global:

* ``third_party/qvm/qvm/quasi_distr.py:3,7-10`` pruning (``abs(v) > ACCURACY``) at
  every construction -> ``prune``;
* ``:13-20`` ``from_counts``; ``:45-53`` ``split``; ``:55-60`` ``merge`` (XOR of keys,
  product of values, *assignment* on collision); ``:62-86`` ``+ - *``;
  ``:28-43`` ``nearest_probability_distribution``;
* ``third_party/qvm/qvm/virtual_gates.py:105-124`` (move), ``:179-194`` (cz/cx/cy),
  ``:262-286`` (rzz/cp) knit formulas, evaluated left to right exactly as written;
* ``third_party/qvm/qvm/virtual_circuit.py:133-171,216-228`` (broadcast of fragment
  results to global labels, per-label merge) and ``:50-68,193-194`` (level loop,
  last virtual gate first, chunks of ``n_k``).

``acc = 1e-5`` is the reference's behaviour, ``acc = 0`` the exact mode.
Pinned against the reference's own code through ``tests/golden/knit_cases.json``.
"""
import itertools
from math import cos, sin

RZZ_EPS = 1e-5


def prune(d, acc):
    return {k: v for k, v in d.items() if abs(v) > acc}


def from_counts(counts, acc):
    shots = sum(counts.values())
    return prune({int("".join(key.split()), 2): value / shots for key, value in counts.items()}, acc)


def split(d, bit, acc):
    mask = 1 << bit
    lo, hi = {}, {}
    for key, value in d.items():
        if key & mask:
            hi[key & ~mask] = value
        else:
            lo[key] = value
    return prune(lo, acc), prune(hi, acc)


def merge(a, b, acc):
    out = {}
    for k1, v1 in a.items():
        for k2, v2 in b.items():
            out[k1 ^ k2] = v1 * v2
    return prune(out, acc)


def add(a, b, acc):
    out = {k: a[k] + b.get(k, 0.0) for k in a}
    out.update({k: b[k] for k in b if k not in a})
    return prune(out, acc)


def sub(a, b, acc):
    out = {k: a[k] - b.get(k, 0.0) for k in a}
    out.update({k: -b[k] for k in b if k not in a})
    return prune(out, acc)


def scale(a, s, acc):
    return prune({k: v * s for k, v in a.items()}, acc)


def nearest_probability_distribution(d):
    ordered = sorted(d.items(), key=lambda kv: kv[1])
    num = len(ordered)
    beta = 0.0
    out = {}
    for key, val in ordered:
        if val + beta / num < 0:
            beta += val
            num -= 1
        else:
            out[key] = val + beta / num
    return out


# ---------------------------------------------------------------- per-gate knit
_CHAIN_SIGNS = {"move": (+1, +1, +1, -1, +1, -1, +1, -1),
                "cz": (+1, +1, +1, -1, +1, -1),
                "cx": (+1, +1, +1, -1, +1, -1),
                "cy": (+1, +1, +1, -1, +1, -1)}


def knit_gate(kind, m_theta, results, clbit, acc):
    if kind in _CHAIN_SIGNS:
        signs = _CHAIN_SIGNS[kind]
        assert len(results) == len(signs)
        total = None
        for r, sg in zip(results, signs):
            r0, r1 = split(r, clbit, acc)
            diff = sub(r0, r1, acc)
            if total is None:
                total = diff
            elif sg > 0:
                total = add(total, diff, acc)
            else:
                total = sub(total, diff, acc)
        return scale(total, 0.5, acc)
    # rzz family
    c, s = cos(m_theta / 2), sin(m_theta / 2)
    if abs(c) < RZZ_EPS:
        r, _ = split(results[0], clbit, acc)
        return scale(r, s ** 2, acc)
    if abs(s) < RZZ_EPS:
        r, _ = split(results[0], clbit, acc)
        return scale(r, c ** 2, acc)
    r0, _ = split(results[0], clbit, acc)
    r1, _ = split(results[1], clbit, acc)
    r23 = add(results[2], results[3], acc)
    r45 = add(results[4], results[5], acc)
    r230, r231 = split(r23, clbit, acc)
    r450, r451 = split(r45, clbit, acc)
    mixed = add(sub(sub(r230, r231, acc), r450, acc), r451, acc)
    # "(...) * cos(m/2) * sin(m/2)" is two successive scalar multiplications (each prunes)
    return add(add(scale(r0, c ** 2, acc), scale(r1, s ** 2, acc), acc),
               scale(scale(mixed, c, acc), s, acc), acc)


# ---------------------------------------------------------------- driver
def knit(frag_results, frag_touches, vgates, n_clbits, acc):
    """``frag_results``: list (fragment order) of lists of dicts in fragment-label order;
    ``frag_touches[f][k]``: does vgate k touch fragment f; ``vgates``: [(kind, m_theta, n_k)]."""
    K = len(vgates)
    radices = [v[2] for v in vgates]
    global_labels = list(itertools.product(*[range(r) for r in radices])) if K else [()]
    lists = []
    for res, touch in zip(frag_results, frag_touches):
        per = [tuple(range(radices[k])) if touch[k] else (-1,) for k in range(K)]
        labels = list(itertools.product(*per)) if K else [()]
        by_label = dict(zip(labels, res))
        lists.append([by_label[tuple(g[k] if touch[k] else -1 for k in range(K))] for g in global_labels])
    merged = []
    for group in zip(*lists):
        m = group[0]
        for other in group[1:]:
            m = merge(m, other, acc)
        merged.append(m)
    if K == 0:
        return merged[0]
    clbit = n_clbits + K - 1
    for k in reversed(range(K)):
        kind, m_theta, n_k = vgates[k]
        merged = [knit_gate(kind, m_theta, merged[i:i + n_k], clbit, acc)
                  for i in range(0, len(merged), n_k)]
        clbit -= 1
    return merged[0]
