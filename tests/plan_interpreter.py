"""Numpy interpreter of compiled device programs (test tool, CPU only).

Executes exactly the data ``qck_sim_fragments`` consumes - the op records, the
sweeps with their tile positions, the matrix pool, the label digits and the
fold masks - the way ``csrc/sim.cu`` does, so that the host compiler
(``compiler.py``) can be checked against the oracle without a GPU.  It is not a
fallback: the product never imports it.
"""
import numpy as np

from hardwareawareoptimalquantumcircuitcuttingandknitting_b200 import _lib


def _bits(idx, q):
    return (idx >> q) & 1


def apply_op(psi, idx, kind, g0, g1, mats, moff):
    """One op on bit positions g0 / g1 of the index space ``idx`` (whole state or one tile)."""
    if kind == _lib.OP_U1:
        m = mats[moff:moff + 8].view(np.complex128).reshape(2, 2)
        lo = psi[(_bits(idx, g0) == 0)]
        hi = psi[(_bits(idx, g0) == 1)]
        new = psi.copy()
        new[_bits(idx, g0) == 0] = m[0, 0] * lo + m[0, 1] * hi
        new[_bits(idx, g0) == 1] = m[1, 0] * lo + m[1, 1] * hi
        return new
    if kind == _lib.OP_CX:
        src = np.where(_bits(idx, g0) == 1, idx ^ (1 << g1), idx)
        return psi[src]
    if kind == _lib.OP_CZ:
        return np.where((_bits(idx, g0) & _bits(idx, g1)) == 1, -psi, psi)
    m = mats[moff:moff + 32].view(np.complex128).reshape(4, 4)
    sub = _bits(idx, g0) + 2 * _bits(idx, g1)
    base = idx & ~((1 << g0) | (1 << g1))
    new = np.zeros_like(psi)
    for r in range(4):
        for c in range(4):
            srcidx = base | ((c & 1) << g0) | ((c >> 1) << g1)
            new += np.where(sub == r, m[r, c] * psi[srcidx], 0)
    return new


def cond_matrix(mats, terms, base):
    """Tile-resolved matrix of an OP_U1X chain: ordered product of the term matrices the bits of
    ``base`` select (first term applied first)."""
    m = np.eye(2, dtype=np.complex128)
    for t in terms:
        e0, off = int(t[1]), int(t[3])
        pair = mats[off:off + 16].view(np.complex128).reshape(2, 2, 2)
        m = pair[(base >> e0) & 1] @ m
    return m


def phase_scalar(mats, terms, base):
    s = 1.0 + 0j
    for t in terms:
        e0, e1, off, n = int(t[1]), int(t[2]), int(t[3]), int(t[5])
        sc = mats[off:off + 2 * n].view(np.complex128)
        s *= sc[((base >> e0) & 1) + (2 * ((base >> e1) & 1) if e1 >= 0 else 0)]
    return s


def apply_u1_matrix(psi, idx, g0, m, where=None):
    lo_sel = _bits(idx, g0) == 0
    hi_sel = _bits(idx, g0) == 1
    partner_hi = psi[idx | (1 << g0)]
    partner_lo = psi[idx & ~(1 << g0)]
    new = np.where(lo_sel, m[0, 0] * psi + m[0, 1] * partner_hi, m[1, 0] * partner_lo + m[1, 1] * psi)
    return new if where is None else np.where(where, new, psi)


def run_plan(program, plan, label, return_state=False):
    """-> output row (1-D float64) of one instance (or the final statevector)."""
    mats = program.mats
    digits = list(np.unravel_index(int(label), program.radix)) if program.radix else []
    N = plan.n_state
    psi = np.zeros(1 << N, dtype=np.complex128)
    psi[0] = 1.0
    idx = np.arange(1 << N)
    for si, (positions, b, e) in enumerate(plan.sweeps):
        cluster_pos, members_left = None, 0
        seg = plan.ops[b:e]
        if getattr(plan, "shared_prefix", False):
            assert len(plan.sweeps) == 2 and len(positions) == N
            if si == 0:       # runs once for all instances: nothing in it may depend on the label
                assert all(int(op[4]) < 0 for op in seg if int(op[0]) != _lib.OP_CLUSTER), \
                    "label-dependent op inside a shared prefix"       # (cluster headers keep a position there)
        tile_mask = sum(1 << p for p in positions)
        skip = 0
        for oi, op in enumerate(seg):
            kind, q0, q1, mat, sel, stride, n_live, _ = (int(x) for x in op)
            if skip:
                skip -= 1
                assert kind == _lib.OP_TERM
                continue
            if kind in (_lib.OP_U1X, _lib.OP_PHASE):
                # tile-resolved forms: the outside (diag) qubits select a variant per tile
                n_terms = q1
                terms = seg[oi + 1:oi + 1 + n_terms]
                skip = n_terms
                outside = set()
                for t in terms:
                    assert int(t[0]) == _lib.OP_TERM
                    for e in (int(t[1]), int(t[2])):
                        if e >= 0:
                            assert not (tile_mask >> e) & 1, "term qubit must be outside the tile"
                            outside.add(e)
                outside = sorted(outside)
                for combo in range(1 << len(outside)):
                    base = sum(((combo >> j) & 1) << e for j, e in enumerate(outside))
                    sel_amp = np.ones(len(idx), dtype=bool)
                    for j, e in enumerate(outside):
                        sel_amp &= _bits(idx, e) == ((combo >> j) & 1)
                    if kind == _lib.OP_PHASE:
                        psi = np.where(sel_amp, phase_scalar(mats, terms, base) * psi, psi)
                    else:
                        psi = apply_u1_matrix(psi, idx, positions[q0], cond_matrix(mats, terms, base), sel_amp)
                continue
            if kind == _lib.OP_CLUSTER:           # header: members address the 3 positions by rank
                cluster_pos, members_left = [mat, sel, stride], q0
                assert cluster_pos == sorted(cluster_pos) and len(set(cluster_pos)) == 3
                live = n_live if 0 < n_live <= len(positions) else len(positions)
                assert all(0 <= p < live for p in cluster_pos)
                cluster_live = n_live
                if n_live > 0 and len(plan.sweeps) == 1:
                    assert np.all(psi[(1 << n_live):] == 0), "cluster n_live hides populated amplitudes"
                continue
            if members_left > 0:
                members_left -= 1
                n_live = cluster_live          # the device uses the header's n_live for every member
                q0 = cluster_pos[q0]
                if kind != _lib.OP_U1:
                    q1 = cluster_pos[q1]
            moff = mat + (digits[sel] * stride if sel >= 0 else 0)
            g0 = positions[q0]
            g1 = positions[q1] if kind != _lib.OP_U1 else 0
            psi = apply_op(psi, idx, kind, g0, g1, mats, moff)
            if n_live > 0 and len(plan.sweeps) == 1:
                assert np.all(psi[(1 << n_live):] == 0), "amplitudes beyond n_live must stay zero"
    if return_state:
        return psi
    prob = psi.real ** 2 + psi.imag ** 2
    n_out = len(plan.out_pos)
    row = np.zeros(1 << n_out)
    for o in range(1 << n_out):
        base, dead = 0, False
        for j in range(n_out):
            if (o >> j) & 1:
                if plan.out_pos[j] < 0:
                    dead = True
                    break
                base |= 1 << plan.out_pos[j]
        if dead:
            continue
        sub, acc = 0, 0.0
        while True:
            sign = -1.0 if bin(sub & plan.sign_mask).count("1") & 1 else 1.0
            acc += sign * prob[base | sub]
            sub = (sub - plan.sum_mask) & plan.sum_mask
            if sub == 0:
                break
        row[o] = acc
    return row


def run_program(program, fold=True, labels=None):
    """-> [num_labels, row_len] table like FragmentExecutor.run."""
    out = np.zeros((program.num_labels, program.row_len(fold)))
    for plan in program.plans(fold):
        for label in plan.labels:
            if labels is not None and label not in labels:
                continue
            out[label] = run_plan(program, plan, label)
    return out


def run_program_deduped(program, fold=True):
    """What FragmentExecutor.run does with instance de-duplication: only the representatives of
    ``program.canonical_labels()`` are simulated, every other row is a copy of its representative's."""
    src = program.canonical_labels()
    reps = set(int(x) for x in np.unique(src))
    out = run_program(program, fold, labels=reps)
    return out[src]


def run_plan_warp(program, plan, label, dfs_levels=2):
    """The register-resident kernel with its outcome tree split over warps: when no outcome bit is a column
    bit and the plan has more than ``dfs_levels`` branch points, the first ones are FORCED per work item
    (2^split items per instance), every item accumulates its leaves, and the partial rows are added in item
    order (csrc/sim_warp_kernel.inc: the last warp to arrive does that)."""
    nb = plan.warp_base
    n_anc = plan.n_state - nb
    acc_mode = all(p < nb for p in plan.out_pos)
    split = n_anc - dfs_levels if (acc_mode and n_anc > dfs_levels) else 0
    row = np.zeros(1 << len(plan.out_pos))
    for forced in range(1 << split):
        row = row + _run_plan_warp_item(program, plan, label, split, forced)
    return row


def _run_plan_warp_item(program, plan, label, split, forced):
    """One work item of the register-resident kernel (csrc/sim_warp_kernel.inc) step for step: the state holds only the
    ``plan.warp_base`` fragment qubits, a CX onto a state bit >= warp_base is a branch point walked depth
    first (stash, project onto 0, run to the end, fold the leaf, restore, project onto 1), the leaf fold
    takes the outcome bits as the virtual state bits above warp_base.  Must equal ``run_plan`` (the same
    ops on the ancilla-enlarged state) bit for bit in exact arithmetic, to rounding in floating point."""
    nb = plan.warp_base
    assert nb > 0 and len(plan.sweeps) == 1
    mats = program.mats
    digits = list(np.unravel_index(int(label), program.radix)) if program.radix else []
    idx = np.arange(1 << nb)
    psi = np.zeros(1 << nb, dtype=np.complex128)
    psi[0] = 1.0
    _, b, e = plan.sweeps[0]
    ops = plan.ops[b:e]
    n_out = len(plan.out_pos)
    row = np.zeros(1 << n_out)
    base_mask = (1 << nb) - 1
    base_sum, base_sign = plan.sum_mask & base_mask, plan.sign_mask & base_mask
    free = [(j, p) for j, p in enumerate(plan.out_pos) if 0 <= p < nb]
    pc, outcomes, stack = 0, 0, []                 # stack: (return pc, saved state, qubit, outcome bit)
    while True:
        if pc < len(ops):
            kind, q0, q1, mat, sel, stride, _, _ = (int(x) for x in ops[pc])
            if kind == _lib.OP_U1:
                moff = mat + (digits[sel] * stride if sel >= 0 else 0)
                psi = apply_op(psi, idx, kind, q0, 0, mats, moff)
            elif kind == _lib.OP_CX and q1 >= nb:  # branch point
                k = q1 - nb
                keep = 0
                if k < split:                      # forced by the work item: no walk
                    keep = (forced >> k) & 1
                else:
                    stack.append((pc, psi.copy(), q0, k))
                outcomes = (outcomes & ~(1 << k)) | (keep << k)
                psi = np.where(_bits(idx, q0) == keep, psi, 0.0)
            elif kind in (_lib.OP_CX, _lib.OP_CZ):
                assert q0 < nb and q1 < nb
                psi = apply_op(psi, idx, kind, q0, q1, mats, 0)
            else:
                raise AssertionError(f"op kind {kind} in a register-resident program")
            pc += 1
            continue
        # leaf
        prob = psi.real ** 2 + psi.imag ** 2
        P = np.where(np.array([bin(int(i) & base_sign).count("1") & 1 for i in idx]) == 1, -prob, prob)
        virt = outcomes << nb
        leaf_sign = -1.0 if bin(virt & plan.sign_mask).count("1") & 1 else 1.0
        col_fixed = 0
        for j, p in enumerate(plan.out_pos):
            if p >= nb and (virt >> p) & 1:
                col_fixed |= 1 << j
        for o in range(1 << len(free)):
            base, col = 0, col_fixed
            for r, (j, p) in enumerate(free):
                if (o >> r) & 1:
                    base |= 1 << p
                    col |= 1 << j
            sub, acc = 0, 0.0
            while True:
                acc += P[base | sub]
                sub = (sub - base_sum) & base_sum
                if sub == 0:
                    break
            row[col] += leaf_sign * acc
        # backtrack
        resumed = False
        while stack:
            rpc, saved, q, bit = stack[-1]
            if (outcomes >> bit) & 1:
                stack.pop()
                continue
            outcomes |= 1 << bit
            psi = np.where(_bits(idx, q) == 1, saved, 0.0)
            pc = rpc + 1
            resumed = True
            break
        if not resumed:
            break
    return row


def run_tree(program, label_range=None):
    """The tree-walk simulation (csrc/sim_tree_kernel.inc) step for step: level kernels over (parent node,
    choice) items, partial rows of the leaves, then the per-(label, column) combine.  -> [num_labels, row_len]"""
    tree = program.tree()
    assert tree is not None
    mats = program.mats
    nb = tree.n_base
    idx = np.arange(1 << nb)

    def u1(psi, q, off):
        return apply_op(psi, idx, _lib.OP_U1, q, 0, mats, off)

    def segment(psi, seg):
        for row in tree.ops[seg[0]:seg[1]]:
            kind, q0, q1, mat = (int(x) for x in row[:4])
            psi = apply_op(psi, idx, kind, q0, q1, mats, mat)
        return psi

    states = None
    n_levels = len(tree.levels)
    part = None
    for lv, L in enumerate(tree.levels):
        n_items = tree.node_counts[lv + 1]
        B = len(L.choices)
        new_states = np.zeros((n_items, 1 << nb), dtype=np.complex128)
        for it in range(n_items):
            parent, c = divmod(it, B)
            if lv == 0:
                psi = np.zeros(1 << nb, dtype=np.complex128)
                psi[0] = 1.0
                psi = segment(psi, tree.seg0)
            else:
                psi = states[parent]
            v, oc = L.choices[c]
            if L.pre_off >= 0:
                psi = u1(psi, L.qubit, L.pre_off + 8 * v)
            if oc >= 0:
                psi = np.where(_bits(idx, L.qubit) == oc, psi, 0.0)
            if L.post_off >= 0:
                psi = u1(psi, L.qubit, L.post_off + 8 * v)
            new_states[it] = segment(psi, L.seg)
        states = new_states
    # leaves -> partial rows
    n_free = len(tree.free)
    part = np.zeros((tree.node_counts[-1], 1 << n_free))
    for it in range(tree.node_counts[-1]):
        node, parity, base_sign = it, 0, 0
        for L in reversed(tree.levels):
            node, cl = divmod(node, len(L.choices))
            v, oc = L.choices[cl]
            if L.kind == _lib.TREE_SLOT and oc > 0:
                parity ^= 1
            elif L.kind == _lib.TREE_TERMINAL and L.meas[v]:
                base_sign |= 1 << L.qubit
        prob = states[it].real ** 2 + states[it].imag ** 2
        sgn = np.array([(-1.0) ** (bin(int(i) & base_sign).count("1") + parity) for i in idx])
        P = sgn * prob
        for o in range(1 << n_free):
            base = sum(1 << p for r, (_, p) in enumerate(tree.free) if (o >> r) & 1)
            sub, acc = 0, 0.0
            while True:
                acc += P[base | sub]
                sub = (sub - tree.base_sum) & tree.base_sum
                if sub == 0:
                    break
            part[it, o] = acc
    # combine: one (label, column) at a time
    out = np.zeros((program.num_labels, 1 << tree.n_out_bits))
    labels = range(program.num_labels) if label_range is None else range(*label_range)
    for label in labels:
        digits = list(np.unravel_index(label, program.radix)) if program.radix else []
        for col in range(1 << tree.n_out_bits):
            o = sum(1 << r for r, (j, _) in enumerate(tree.free) if (col >> j) & 1)
            base, fork = [], []
            for L in tree.levels:
                if L.kind == _lib.TREE_MMEAS:
                    base.append((col >> L.col_bit) & 1)
                    fork.append(0)
                else:
                    rep = L.canon[int(digits[L.digit])]
                    base.append([c for c, (v, _) in enumerate(L.choices) if v == rep][0])
                    fork.append(1 if (L.kind == _lib.TREE_SLOT and L.meas[rep]) else 0)
            acc = 0.0
            for combo in range(1 << sum(fork)):
                node, k = 0, 0
                for L, b, f in zip(tree.levels, base, fork):
                    c = b
                    if f:
                        c += (combo >> k) & 1
                        k += 1
                    node = node * len(L.choices) + c
                acc += part[node, o]
            out[label, col] = acc
    return out
