"""Numpy emulation of csrc/npd.cu (the sort-free, host-round-trip-free nearest_probability_distribution):
the same statistics, key ranges, bins, integer quantisation and tails, step for step, so that the LOGIC of
the kernels is pinned on CPU against the oracle (``oracle.dense.nearest_probability_distribution`` =
quasi_distr.py:28-43) before any GPU time is spent.  ``shards > 1`` emulates the multi-rank form
(qck_npd_stage with the statistics / bins summed across ranks between stage and tail).  ``bin_bits = 11`` is the
geometry of the one-cluster kernel for vectors of at most 2^16 entries (npd_cluster_kernel: 2048 bins per level,
six levels; otherwise the same definitions)."""
import math

import numpy as np

BINS, BIN_BITS, LEVELS = 8192, 13, 5
SEARCH, IDENTITY, SOLVED, NEGATIVE_TOTAL, LOCATED = 0, 1, 2, 3, 4
_M64 = (1 << 64) - 1


def key_of(v):
    b = np.asarray(v, dtype=np.float64).view(np.int64).astype(object)
    return np.array([(0x8000000000000000 - (int(x) & _M64)) if x < 0 else int(x) for x in b.ravel()],
                    dtype=object).reshape(np.shape(v))


def key_of_fast(v):
    b = np.asarray(v, dtype=np.float64).view(np.int64)
    neg = b < 0
    out = b.copy()
    # 0x8000000000000000 - (unsigned)b  for negative b, as signed 64-bit
    out[neg] = (np.uint64(0x8000000000000000) - b[neg].view(np.uint64)).view(np.int64)
    return out


def val_of(k):
    k = int(k)
    b = (0x8000000000000000 - (k & _M64)) & _M64 if k < 0 else k
    return float(np.array([b], dtype=np.uint64).view(np.float64)[0])


class State:
    pass


def _set_level(s, BIN_BITS=BIN_BITS):
    width = (s.hi - s.lo)
    bits = 0 if width <= 1 else (width - 1).bit_length()
    s.shift = bits - BIN_BITS if bits > BIN_BITS else 0
    w = val_of(s.hi) - val_of(s.lo + 1)
    cnt = max(1, s.sel_cnt)
    cbits = cnt.bit_length()
    e = 0
    if w > 0.0 and math.isfinite(w):
        e = 61 - cbits - (math.frexp(w)[1] - 1 + 1)      # ilogb(w) = frexp exponent - 1
    s.qexp = e


def _plan(s, BIN_BITS=BIN_BITS, lo_zero=False):
    s.level, s.under_sum, s.under_cnt, s.beta, s.num, s.shift_val, s.t0 = 0, 0.0, 0.0, 0.0, s.alive, 0.0, -math.inf
    if not (s.alive > 0.0) or not (s.vmin < 0.0):
        s.status = IDENTITY
    elif s.sum < 0.0:
        s.status = NEGATIVE_TOTAL
    else:
        s.status = SEARCH
        t_ub = -s.neg_sum * (1.0 + 1e-9)
        # (the cluster kernel starts above zero: G(0) = neg_sum < 0, so t0 > 0 and every entry <= 0 is dropped)
        s.lo = 0 if lo_zero else int(key_of_fast(np.array([s.vmin]))[0]) - 1
        s.hi = int(key_of_fast(np.array([t_ub]))[0])
        s.sel_cnt = int(s.alive)
        _set_level(s, BIN_BITS)


def _select(s, bin_cnt, bin_q, BIN_BITS=BIN_BITS):
    BINS = 1 << BIN_BITS
    if s.status == LOCATED:
        s.beta = s.under_sum
        s.num = s.alive - s.under_cnt
        s.shift_val = s.beta / s.num if s.num > 0 else 0.0
        s.t0 = val_of(s.lo + 1)
        s.status = SOLVED
        return
    if s.status != SEARCH:
        return
    lo, hi, shift = s.lo, s.hi, s.shift
    width = hi - lo
    lo_val = val_of(lo + 1)
    rest, under = s.alive - s.under_cnt, s.under_sum
    c_ex, q_ex, found = 0, 0, None
    for j in range(BINS):
        c_ex += int(bin_cnt[j])
        q_ex += int(bin_q[j])
        off = min((j + 1) << shift, width)
        ub = val_of(lo + off)
        g = under + (float(c_ex) * lo_val + math.ldexp(float(q_ex), -s.qexp)) + ub * (rest - float(c_ex))
        if g >= 0.0 or off == width:
            found = j
            break
    j = found
    off_lo, off_hi = j << shift, min((j + 1) << shift, width)
    cnt_j = int(bin_cnt[j])
    s.lo, s.hi, s.sel_cnt = lo + off_lo, lo + off_hi, cnt_j
    s.level += 1
    if cnt_j == 0 or off_hi - off_lo <= 1:
        s.status = LOCATED
    else:
        _set_level(s, BIN_BITS)


def nearest_probability_distribution(p, acc=0.0, shards=1, bin_bits=BIN_BITS):
    """-> (result, state).  Mirrors qck_npd_async (shards == 1) / the staged multi-rank flow."""
    p = np.asarray(p, dtype=np.float64).copy()
    parts = np.array_split(np.arange(len(p)), shards)
    alive = np.abs(p) > acc
    keys = key_of_fast(p)
    s = State()
    # stage STATS per shard, reduced
    s.sum = sum(float(p[i][alive[i]].sum()) for i in parts)
    s.vmin = min([float(p[i][alive[i]].min()) if alive[i].any() else math.inf for i in parts])
    s.neg_sum = sum(float(p[i][alive[i] & (p[i] < 0)].sum()) for i in parts)
    s.alive = float(alive.sum())
    s.neg_cnt = float((alive & (p < 0)).sum())
    BINS, levels = 1 << bin_bits, -(-64 // bin_bits)
    s.levels = levels
    _plan(s, bin_bits, lo_zero=bin_bits == 11)
    passes = 0
    for _ in range(levels + 1):
        if s.status not in (SEARCH, LOCATED):
            break
        passes += 1
        bin_cnt = np.zeros(BINS, dtype=object)
        bin_q = np.zeros(BINS, dtype=object)
        us, uc = 0.0, 0.0
        for i in parts:
            a, k, v = alive[i], keys[i], p[i]
            below = a & (k <= s.lo)
            us += float(v[below].sum())
            uc += float(below.sum())
            if s.status == SEARCH:
                inr = a & (k > s.lo) & (k <= s.hi)
                kk = k[inr].astype(object)
                b = np.array([(int(x) - s.lo - 1) >> s.shift for x in kk], dtype=np.int64)
                lo_val = val_of(s.lo + 1)
                q = [int(np.rint(math.ldexp(float(x) - lo_val, s.qexp))) for x in v[inr]]
                for bb, qq in zip(b.tolist(), q):
                    assert 0 <= bb < BINS and qq >= 0
                    bin_cnt[bb] += 1
                    bin_q[bb] += qq
        assert all(int(x) < (1 << 62) for x in bin_q)
        s.under_sum, s.under_cnt = us, uc
        _select(s, bin_cnt, bin_q, bin_bits)
    s.passes = passes
    if s.status == SOLVED:
        keep = alive & (keys > s.lo)
        p = np.where(keep, p + s.shift_val, 0.0)
    elif s.status == IDENTITY and acc > 0.0:
        p = np.where(alive, p, 0.0)
    return p, s
