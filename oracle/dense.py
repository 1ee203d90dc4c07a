"""Dense numpy forms of the knit algebra (TEST INFRASTRUCTURE - see oracle/__init__.py).

* ``pext`` / ``pdep``: the bit-index maps between a fragment's compact output
  index and the full-circuit bitstring (bit i = clbit i).  These are the
  "bit-exact integer work" of the path (reference: XOR of disjoint-support keys
  in ``quasi_distr.py:55-60``).
* ``signed_fold``: ``q_f = sum_bits (-1)^bits p_f`` over the config bits
  (closed form of ``split`` + the signed sums of ``virtual_gates.py:105-124,
  179-194,262-286``).
* ``contract``: ``P(x) = sum_l prod_k w_k(l_k) prod_f q_f(x_f | l)``.
* ``knit_outer``: the K = 0 case, ``out[y] = prod_f p_f[pext(y, mask_f)]``.
* ``nearest_probability_distribution`` (``quasi_distr.py:28-43``) on a dense
  vector, with the reference's element count (= entries that survive pruning).
* ``hellinger_fidelity``: qiskit-terra 0.25.2.1
  ``quantum_info/analysis/distance.py`` (third-party, not vendored): both inputs
  normalised by their own totals, ``H^2 = 1/2 sum (sqrt p - sqrt q)^2`` over the
  union of keys, ``F = (1 - H^2)^2`` (call site ``src/HwAwareCutter/Utilities.py:224``).
"""
import itertools
import math

import numpy as np


def pext(y, mask):
    """Gather the bits of ``y`` selected by ``mask`` into a compact integer (array ok)."""
    y = np.asarray(y, dtype=np.uint64)
    out = np.zeros_like(y)
    j = 0
    for b in range(64):
        if (mask >> b) & 1:
            out |= ((y >> np.uint64(b)) & np.uint64(1)) << np.uint64(j)
            j += 1
    return out


def pdep(x, mask):
    x = np.asarray(x, dtype=np.uint64)
    out = np.zeros_like(x)
    j = 0
    for b in range(64):
        if (mask >> b) & 1:
            out |= ((x >> np.uint64(j)) & np.uint64(1)) << np.uint64(b)
            j += 1
    return out


def knit_outer(tables, masks, y_begin, y_end):
    y = np.arange(y_begin, y_end, dtype=np.uint64)
    out = np.ones(y.shape[0])
    for t, m in zip(tables, masks):
        out = out * np.asarray(t)[pext(y, m).astype(np.int64)]
    return out


def signed_fold(dist, n_clbits, K, out_mask):
    """dict over (n_clbits + K)-bit keys -> dense row over popcount(out_mask) bits,
    entries with an odd number of config bits set counted negatively."""
    m = bin(out_mask).count("1")
    row = np.zeros(1 << m)
    low = (1 << n_clbits) - 1
    for key, v in dist.items():
        x = key & low
        assert x & ~out_mask == 0, "fragment wrote a clbit outside its output mask"
        sign = -1.0 if bin(key >> n_clbits).count("1") & 1 else 1.0
        row[int(pext(np.uint64(x), out_mask))] += sign * v
    return row


def contract(folded, touches, coeffs, out_masks, n_out):
    """``folded[f]``: array [L_f, 2^m_f] in fragment-label order; ``coeffs[k]``: list of the
    bit-0 coefficients a_i of vgate k (the bit-1 coefficient is folded into the sign)."""
    K = len(coeffs)
    radices = [len(c) for c in coeffs]
    out = np.zeros(1 << n_out)
    y = np.arange(1 << n_out, dtype=np.uint64)
    idx = [pext(y, m).astype(np.int64) for m in out_masks]
    strides = []
    for touch in touches:
        s, acc = [0] * K, 1
        for k in reversed(range(K)):
            if touch[k]:
                s[k] = acc
                acc *= radices[k]
        strides.append(s)
    for label in itertools.product(*[range(r) for r in radices]):
        w = 1.0
        for k, d in enumerate(label):
            w *= coeffs[k][d]
        term = np.full(1 << n_out, w)
        for f, rows in enumerate(folded):
            lf = sum(d * strides[f][k] for k, d in enumerate(label))
            term = term * rows[lf][idx[f]]
        out += term
    return out


def nearest_probability_distribution(v, acc=0.0):
    """Dense restatement: entries with |v| <= acc do not exist for the reference
    (they were pruned), so they neither count in ``num`` nor receive the shift."""
    v = np.asarray(v, dtype=np.float64)
    alive = np.abs(v) > acc
    order = np.argsort(v[alive], kind="stable")
    keys = np.nonzero(alive)[0][order]
    vals = v[keys]
    num = len(vals)
    beta = 0.0
    out = np.zeros_like(v)
    for k, val in zip(keys.tolist(), vals.tolist()):
        if val + beta / num < 0:
            beta += val
            num -= 1
        else:
            out[k] = val + beta / num
    return out


def hellinger_fidelity(p, q):
    """dict or dense inputs."""
    if not isinstance(p, dict):
        p = {i: x for i, x in enumerate(np.asarray(p).tolist()) if x != 0.0}
    if not isinstance(q, dict):
        q = {i: x for i, x in enumerate(np.asarray(q).tolist()) if x != 0.0}
    sp, sq = sum(p.values()), sum(q.values())
    total = 0.0
    for k in set(p) | set(q):
        a = p.get(k, 0.0) / sp if k in p else 0.0
        b = q.get(k, 0.0) / sq if k in q else 0.0
        total += (math.sqrt(a) - math.sqrt(b)) ** 2
    dist = math.sqrt(total) / math.sqrt(2)
    return (1 - dist ** 2) ** 2


def hellinger_fidelity_dense(p, q):
    p = np.asarray(p, dtype=np.float64)
    q = np.asarray(q, dtype=np.float64)
    h2 = 0.5 * np.sum((np.sqrt(p / p.sum()) - np.sqrt(q / q.sum())) ** 2)
    return float((1 - h2) ** 2)
