"""Index arithmetic of the knit kernels restated in plain Python (bit-exact work, no GPU):

* knit_outer recomputes its per-thread factors on a row change from the chunk-local output index
  y_lo = (e << 10) | (tid << 2) | j (KO_THREADS = 256, KO_ITEMS = 4, KO_CHUNK_BITS = 12; csrc/knit.cu).  The table
  index of a fragment is pext(y_lo, lo_mask); the kernel takes it apart over the three disjoint bit ranges - one
  8-bit pext of tid per fragment, closed forms for the two-bit parts of j and e - instead of one software pext per
  entry.  The split must be the pext for EVERY mask.
* contract_scatter_kernel sends tile entry (i, j) to output pdep(i, maskA) | pdep(j, maskB): with masks that are
  disjoint and cover the output (virtual_circuit.py: every clbit belongs to exactly one fragment) this is a
  bijection, the inverse of (pext(y, maskA), pext(y, maskB)) that the reference's key arithmetic
  (quasi_distr.py:55-60: merged[k1 ^ k2]) implies."""
import random


def pext(x: int, mask: int) -> int:
    out, j = 0, 0
    while mask:
        low = mask & -mask
        if x & low:
            out |= 1 << j
        j += 1
        mask ^= low
    return out


def pdep(x: int, mask: int) -> int:
    out, j = 0, 0
    while mask:
        low = mask & -mask
        if (x >> j) & 1:
            out |= low
        j += 1
        mask ^= low
    return out


def pext2(x: int, m: int) -> int:          # x, m < 4 (the kernel's closed form)
    return x if m == 3 else (x & 1 if m == 1 else (x >> 1 if m == 2 else 0))


def test_knit_outer_row_change_index_split_is_the_pext():
    rng = random.Random(7)
    masks = [0, 0xFFF, 0x001, 0x800, 0x3FC, 0xC03, 0x555, 0xAAA] + [rng.randrange(1 << 12) for _ in range(200)]
    for m in masks:
        mj, mt, me = m & 3, (m >> 2) & 0xFF, (m >> 10) & 3
        sj = bin(mj).count("1")
        se = sj + bin(mt).count("1")
        for tid in list(range(0, 256, 17)) + [255]:
            tpart = pext(tid, mt) << sj
            for e in range(4):
                for j in range(4):
                    y_lo = 4 * (e * 256 + tid) + j
                    assert y_lo == (e << 10) | (tid << 2) | j
                    assert (pext2(j, mj) | tpart | (pext2(e, me) << se)) == pext(y_lo, m), (hex(m), tid, e, j)


def test_contract_scatter_positions_are_a_bijection():
    rng = random.Random(11)
    for n_out in (4, 7, 10):
        full = (1 << n_out) - 1
        for _ in range(6):
            mask_a = rng.randrange(1, full)          # neither empty nor everything
            mask_b = full ^ mask_a
            ma, mb = bin(mask_a).count("1"), bin(mask_b).count("1")
            seen = set()
            for i in range(1 << ma):
                for j in range(1 << mb):
                    y = pdep(i, mask_a) | pdep(j, mask_b)
                    assert pext(y, mask_a) == i and pext(y, mask_b) == j
                    seen.add(y)
            assert len(seen) == 1 << n_out
