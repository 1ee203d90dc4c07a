"""Host logic of the TMA sweep path (csrc/sim.cu: tma_describe) without a GPU.

``qck_debug_tma_describe`` returns, for every sweep of a streaming plan, exactly what the
warp-specialised kernel consumes: the shared-memory bit order, the load / store boxes, which boxes are
zero-filled and which tiles are visited.  A numpy emulation of the kernel's data movement (state buffer
pre-filled with NaN = "never written") then has to reproduce the plain plan interpreter bit for bit and
must never load a NaN: live-qubit tracking may only skip memory that is provably zero.
"""
import ctypes as C
from importlib import import_module

import numpy as np
import pytest

import plan_interpreter as pi

PKG = "hardwareawareoptimalquantumcircuitcuttingandknitting_b200"
_lib = import_module(f"{PKG}._lib")
compiler = import_module(f"{PKG}.compiler")
gen = import_module(f"{PKG}.generators")
vcm = import_module(f"{PKG}.virtual_circuit")
circuit = import_module(f"{PKG}.circuit")


def _pdep(x, mask):
    out, j = 0, 0
    while mask:
        low = mask & -mask
        if (x >> j) & 1:
            out |= low
        j += 1
        mask ^= low
    return out


def _plan_struct(program, plan):
    arr = (_lib.QckSweep * len(plan.sweeps))()
    for i, (positions, b, e) in enumerate(plan.sweeps):
        arr[i].n_tile = len(positions)
        arr[i].op_begin = b
        arr[i].op_end = e
        arr[i].flags = (int(np.isin(plan.ops[b:e, 0], (_lib.OP_U1X, _lib.OP_PHASE)).any())
                        | 2 * int((plan.ops[b:e, 0] == _lib.OP_CLUSTER).any()))
        for j, x in enumerate(positions):
            arr[i].pos[j] = x
    st = _lib.QckSimPlan()
    st.n_state_qubits = plan.n_state
    st.n_sweeps = len(plan.sweeps)
    st.sweeps = arr
    return st, arr


def describe(st, i, live, last, batch=1, n_local=0, rank=0):
    lib = _lib.load()
    geom = (C.c_int32 * 8)()
    perm = (C.c_int32 * 16)()
    ld_off, st_off = (C.c_uint64 * 128)(), (C.c_uint64 * 128)()
    ld_slot, st_slot = (C.c_uint32 * 128)(), (C.c_uint32 * 128)()
    enum_mask, n_work, fixed = C.c_uint64(), C.c_uint64(), C.c_uint64()
    ok = lib.qck_debug_tma_describe(C.byref(st), i, live, int(last), batch, n_local, rank, geom, perm, ld_off, ld_slot,
                                    st_off, st_slot, C.byref(enum_mask), C.byref(n_work), C.byref(fixed))
    if not ok:
        return None
    lowc, h, k, n_load, n_store, zf_shift, zf_mask, n_enum = list(geom)
    return dict(lowc=lowc, h=h, k=k, zf_shift=zf_shift, zf_mask=zf_mask, n_enum=n_enum, perm=list(perm),
                ld=[(ld_off[j], ld_slot[j]) for j in range(n_load)],
                st=[(st_off[j], st_slot[j]) for j in range(n_store)],
                enum_mask=enum_mask.value, n_work=n_work.value, fixed_base=fixed.value)


def emulate(program, plan, label):
    """-> (final state, bytes loaded, bytes stored) of the TMA path for one instance."""
    st, _keep = _plan_struct(program, plan)
    mats = program.mats
    digits = list(np.unravel_index(int(label), program.radix)) if program.radix else []
    N = plan.n_state
    state = np.full(1 << N, np.nan + 1j * np.nan, dtype=np.complex128)
    live = 0
    loaded = stored = 0
    for i, (positions, b, e) in enumerate(plan.sweeps):
        T = len(positions)
        last = i == len(plan.sweeps) - 1
        d = describe(st, i, live, last)
        assert d is not None, f"sweep {i} not eligible"
        lowc, h, k = d["lowc"], d["h"], d["k"]
        perm = d["perm"]
        assert sorted(perm[:T]) == list(range(T))

        def box_offsets(bits):       # box-local index -> amplitude offset in the state
            j = np.arange(1 << bits)
            off = j & ((1 << lowc) - 1)
            for t in range(lowc, bits):
                off = off | (((j >> t) & 1) << (h + t - lowc))
            return off

        ld_box, st_box = box_offsets(d["zf_shift"]), box_offsets(lowc + k)
        assert d["n_work"] == 1 << d["n_enum"]
        tidx = np.arange(1 << T)
        for w in range(d["n_work"]):
            base = _pdep(w, d["enum_mask"])
            tile_live = (base & ~live) == 0
            stage = np.full(1 << T, np.nan + 1j * np.nan, dtype=np.complex128)
            if not tile_live:
                assert last, "dead tiles are only visited by the last sweep"
                stage[:] = 0
            else:
                if live == 0:
                    assert not d["ld"]
                    stage[:] = 0
                    stage[0] = 1
                else:
                    for off, slot in d["ld"]:
                        src = state[(base | off) | ld_box]
                        assert not np.isnan(src).any(), "TMA load would read memory that was never written"
                        stage[slot:slot + len(ld_box)] = src
                        loaded += 16 * len(ld_box)
                    zf = ((tidx >> d["zf_shift"]) & d["zf_mask"]) != 0
                    assert np.isnan(stage[zf]).all(), "zero fill overlaps a loaded box"
                    stage[zf] = 0
                assert not np.isnan(stage).any(), "tile slots neither loaded nor zero-filled"
                cluster_pos, members_left = None, 0
                seg, skip = plan.ops[b:e], 0
                for oi, op in enumerate(seg):
                    kind, q0, q1, mat, sel, stride, n_live, member = (int(x) for x in op)
                    if skip:
                        skip -= 1
                        continue
                    if kind in (_lib.OP_U1X, _lib.OP_PHASE):      # resolved once per tile from its base
                        terms = seg[oi + 1:oi + 1 + q1]
                        skip = q1
                        if kind == _lib.OP_PHASE:
                            stage = pi.phase_scalar(mats, terms, base) * stage
                        else:
                            stage = pi.apply_u1_matrix(stage, tidx, perm[q0], pi.cond_matrix(mats, terms, base))
                        continue
                    if kind == _lib.OP_CLUSTER:
                        cluster_pos, members_left = [perm[mat], perm[sel], perm[stride]], q0
                        continue
                    if members_left > 0:
                        assert member == 1
                        members_left -= 1
                        q0 = cluster_pos[q0]
                        if kind != _lib.OP_U1:
                            q1 = cluster_pos[q1]
                    else:
                        assert member == 0
                        q0 = perm[q0]
                        if kind != _lib.OP_U1:
                            q1 = perm[q1]
                    moff = mat + (digits[sel] * stride if sel >= 0 else 0)
                    stage = pi.apply_op(stage, tidx, kind, q0, q1, mats, moff)
            for off, slot in d["st"]:
                state[(base | off) | st_box] = stage[slot:slot + len(st_box)]
                stored += 16 * len(st_box)
        for p in positions:
            live |= 1 << p
    return state, loaded, stored


def _uncut_program(name, n, depth, **kw):
    circ = gen.gen_circ(name, n, depth, seed=1).decompose_two_qubit()
    virt = vcm.VirtualCircuit(circ)
    (frag,) = virt.active_fragments()
    return compiler.FragmentProgram(virt.fragment_circuits[frag], frag, circ.num_clbits, **kw)


@pytest.mark.parametrize("name,n,depth,onchip,tile", [
    ("syc", 12, 1, 6, 6), ("syc", 12, 3, 6, 6), ("syc", 12, 3, 7, 8), ("hwe", 11, 2, 6, 7), ("qft", 10, 1, 5, 6),
    ("bv", 12, 1, 6, 6), ("aqft", 11, 1, 6, 7), ("syc", 14, 2, 8, 9),
])
def test_tma_path_equals_plain_interpreter(name, n, depth, onchip, tile):
    prog = _uncut_program(name, n, depth, onchip_max=onchip, stream_tile=tile)
    (plan,) = prog.plans()
    assert len(plan.sweeps) > 1
    want = pi.run_plan(prog, plan, 0, return_state=True)
    got, loaded, stored = emulate(prog, plan, 0)
    assert not np.isnan(got).any(), "the last sweep must leave the whole buffer written"
    assert np.abs(got - want).max() < 1e-14
    full = 16 << plan.n_state
    assert stored <= full * len(plan.sweeps) and loaded <= full * (len(plan.sweeps) - 1)


def test_live_tracking_saves_traffic_on_depth_one():
    """A depth-1 circuit expands the state once: about one write of the state in total."""
    prog = _uncut_program("syc", 14, 1, onchip_max=6, stream_tile=7)
    (plan,) = prog.plans()
    want = pi.run_plan(prog, plan, 0, return_state=True)
    got, loaded, stored = emulate(prog, plan, 0)
    assert np.abs(got - want).max() < 1e-14
    full = 16 << plan.n_state
    n_sw = len(plan.sweeps)
    assert n_sw >= 4
    # without tracking: n_sw stores and n_sw - 1 loads of the full state
    assert stored < 2 * full, (stored, full)
    assert loaded < full, (loaded, full)


def test_cut_fragment_with_slots_and_ancillas():
    """Gate-cut fragments (label-selected slot matrices, ancilla bits of mid-circuit measurements) forced
    into the streaming regime."""
    cutting = import_module(f"{PKG}.cutting")
    circ, cut = cutting.make_baseline("syc16d5")
    virt = vcm.VirtualCircuit(cut)
    checked = 0
    for f in virt.active_fragments():
        prog = compiler.FragmentProgram(virt.fragment_circuits[f], f, cut.num_clbits, onchip_max=6, stream_tile=7)
        for plan in prog.plans()[::5]:
            assert len(plan.sweeps) >= 2
            for label in plan.labels[:2]:
                want = pi.run_plan(prog, plan, label, return_state=True)
                got, _, _ = emulate(prog, plan, label)
                assert not np.isnan(got).any()
                assert np.abs(got - want).max() < 1e-14
                checked += 1
    assert checked >= 8


def test_ineligible_sweeps_are_reported():
    prog = _uncut_program("syc", 12, 1, onchip_max=6, stream_tile=6)
    (plan,) = prog.plans()
    st, _keep = _plan_struct(prog, plan)
    assert describe(st, 0, 0, False) is not None
    # live set that does not cover the low run -> the low-run box cannot be loaded
    assert describe(st, 1, 0b1, False) is None


@pytest.mark.parametrize("name,n,depth,onchip,tile,max_sweeps", [
    ("qft", 10, 1, 5, 6, 3), ("aqft", 11, 1, 6, 7, 3), ("qft", 12, 1, 6, 8, 3), ("hwe", 11, 2, 6, 7, 5),
    ("syc", 12, 3, 6, 7, 7),
])
def test_diag_qubits_need_no_tile_residency(name, n, depth, onchip, tile, max_sweeps):
    """Controls of cx and the qubits of cz / cp / rz stay outside the tile (OP_U1X / OP_PHASE, resolved per
    tile): far fewer sweeps for qft-like circuits, same probabilities as the oracle."""
    from oracle import statevector as sv
    prog = _uncut_program(name, n, depth, onchip_max=onchip, stream_tile=tile)
    (plan,) = prog.plans()
    kinds = plan.ops[:, 0].tolist()
    assert len(plan.sweeps) <= max_sweeps
    if name in ("qft", "aqft", "hwe"):
        assert kinds.count(_lib.OP_U1X) > 0 and kinds.count(_lib.OP_TERM) >= kinds.count(_lib.OP_U1X)
    # every term refers to a state bit outside its sweep's tile
    for positions, b, e in plan.sweeps:
        for r in plan.ops[b:e]:
            if r[0] == _lib.OP_TERM:
                assert r[1] not in positions and (r[2] < 0 or r[2] not in positions)
    circ = gen.gen_circ(name, n, depth, seed=1)
    want = sv.dense(sv.exact_distribution(circ), circ.num_clbits)
    assert np.abs(pi.run_plan(prog, plan, 0) - want).max() < 1e-13
    got, _, _ = emulate(prog, plan, 0)
    assert np.abs(got - pi.run_plan(prog, plan, 0, return_state=True)).max() < 1e-14


@pytest.mark.parametrize("name,n,depth,onchip,tile", [("syc", 14, 2, 8, 9), ("hwe", 11, 2, 6, 7), ("qft", 12, 1, 6, 8)])
def test_traffic_accounting_matches_the_emulated_data_movement(name, n, depth, onchip, tile):
    """qck_sim_plan_traffic (the bytes bench.py's roofline figures use) == what the emulated TMA path loads
    and stores; with the fused fold the last sweep stores 8 instead of 16 bytes per amplitude."""
    prog = _uncut_program(name, n, depth, onchip_max=onchip, stream_tile=tile)
    (plan,) = prog.plans()
    st, _keep = _plan_struct(prog, plan)
    lib = _lib.load()
    ld, sd, used = C.c_uint64(), C.c_uint64(), C.c_int()
    assert lib.qck_sim_plan_traffic(C.byref(st), 1, 0, C.byref(ld), C.byref(sd), C.byref(used)) == 0
    _, loaded, stored = emulate(prog, plan, 0)
    assert used.value == 1 and (ld.value, sd.value) == (loaded, stored)
    ld2, sd2 = C.c_uint64(), C.c_uint64()
    assert lib.qck_sim_plan_traffic(C.byref(st), 1, 1, C.byref(ld2), C.byref(sd2), C.byref(used)) == 0
    last_tiles = 1 << (plan.n_state - len(plan.sweeps[-1][0]))
    assert ld2.value == loaded and sd.value - sd2.value == last_tiles * (8 << len(plan.sweeps[-1][0]))
    # three instances move three times as much
    assert lib.qck_sim_plan_traffic(C.byref(st), 3, 0, C.byref(ld2), C.byref(sd2), C.byref(used)) == 0
    assert (ld2.value, sd2.value) == (3 * loaded, 3 * stored)


@pytest.mark.parametrize("seed", range(10))
def test_random_circuits_streaming_schedule_and_tma_layout(seed):
    """Random circuits (every gate kind, cut or uncut) forced into the streaming regime with small tiles:
    the scheduled plans (tile-resolved ops included) reproduce the oracle, and the emulated TMA data
    movement reproduces the plans - for every plan and a few labels of every fragment."""
    import random
    import test_random_circuits_cpu as rc
    from oracle import statevector as sv
    cutting = import_module(f"{PKG}.cutting")
    rng = random.Random(4000 + seed)
    n = rng.randint(6, 8)
    qc = rc.random_circuit(rng, n, rng.randint(20, 40))
    tile = rng.randint(5, 6)
    if seed % 2 == 0:                                    # uncut: one fragment, compare with the oracle
        virt = vcm.VirtualCircuit(qc)
        (f,) = virt.active_fragments()
        prog = compiler.FragmentProgram(virt.fragment_circuits[f], f, qc.num_clbits, onchip_max=4, stream_tile=tile)
        (plan,) = prog.plans()
        assert len(plan.sweeps) >= 2
        want = sv.dense(sv.exact_distribution(qc), n)
        assert np.abs(pi.run_plan(prog, plan, 0) - want).max() < 1e-12
        got, _, _ = emulate(prog, plan, 0)
        assert np.abs(got - pi.run_plan(prog, plan, 0, return_state=True)).max() < 1e-13
    else:                                                # cut: slots, ancillas, several patterns
        cut = cutting.apply_cuts(qc, rc.random_cut(rng, qc, max_gate_cuts=2, wire_cut=(seed % 4 == 1)))
        virt = vcm.VirtualCircuit(cut)
        for f in virt.active_fragments():
            small = compiler.FragmentProgram(virt.fragment_circuits[f], f, cut.num_clbits)
            prog = compiler.FragmentProgram(virt.fragment_circuits[f], f, cut.num_clbits, onchip_max=3, stream_tile=tile)
            ref_rows = pi.run_program(small)
            for plan in prog.plans():
                if len(plan.sweeps) < 2:
                    continue
                for label in plan.labels[:3]:
                    assert np.abs(pi.run_plan(prog, plan, label) - ref_rows[label]).max() < 1e-12
                    try:
                        got, _, _ = emulate(prog, plan, label)
                    except AssertionError as exc:
                        if "not eligible" in str(exc):   # e.g. a low run shorter than 3 bits: plain kernel
                            continue
                        raise
                    assert np.abs(got - pi.run_plan(prog, plan, label, return_state=True)).max() < 1e-13


def emulate_sharded(program, plan, world):
    """The data movement of qck_sim_sweeps_sharded on `world` ranks (shards = separate NaN-filled buffers; a box
    whose rank bits differ from the working rank's is a peer access) -> (concatenated final state, bytes a rank
    moved from / to PEER buffers)."""
    st, _keep = _plan_struct(program, plan)
    mats = program.mats
    N = plan.n_state
    g = world.bit_length() - 1
    n_local = N - g
    shards = [np.full(1 << n_local, np.nan + 1j * np.nan, dtype=np.complex128) for _ in range(world)]
    lmask = (1 << n_local) - 1
    live, peer_bytes = 0, 0
    for i, (positions, b, e) in enumerate(plan.sweeps):
        T = len(positions)
        last = i == len(plan.sweeps) - 1
        written = [np.zeros(1 << n_local, dtype=bool) for _ in range(world)]
        for rank in range(world):
            d = describe(st, i, live, last, n_local=n_local, rank=rank)
            assert d is not None, f"sweep {i} not eligible on rank {rank}"
            lowc, h, k, perm = d["lowc"], d["h"], d["k"], d["perm"]

            def box_offsets(bits):
                j = np.arange(1 << bits)
                off = j & ((1 << lowc) - 1)
                for t in range(lowc, bits):
                    off = off | (((j >> t) & 1) << (h + t - lowc))
                return off

            ld_box, st_box = box_offsets(d["zf_shift"]), box_offsets(lowc + k)
            assert h + k <= n_local, "a box must not span two shards"
            tidx = np.arange(1 << T)
            for w in range(d["n_work"]):
                base = d["fixed_base"] | _pdep(w, d["enum_mask"])
                tile_live = (base & ~live) == 0
                stage = np.full(1 << T, np.nan + 1j * np.nan, dtype=np.complex128)
                if not tile_live:
                    assert last
                    stage[:] = 0
                else:
                    if live == 0:
                        stage[:] = 0
                        stage[0] = 1
                    else:
                        for off, slot in d["ld"]:
                            gidx = base | off
                            src = shards[gidx >> n_local][(gidx & lmask) | ld_box]
                            assert not np.isnan(src).any(), "load of memory that was never written"
                            stage[slot:slot + len(ld_box)] = src
                            peer_bytes += 16 * len(ld_box) * ((gidx >> n_local) != rank)
                        stage[((tidx >> d["zf_shift"]) & d["zf_mask"]) != 0] = 0
                    assert not np.isnan(stage).any()
                    seg, skip = plan.ops[b:e], 0
                    for oi, op in enumerate(seg):
                        kind, q0, q1, mat, sel, stride, n_live, member = (int(x) for x in op)
                        if skip:
                            skip -= 1
                            continue
                        if kind in (_lib.OP_U1X, _lib.OP_PHASE):
                            terms = seg[oi + 1:oi + 1 + q1]
                            skip = q1
                            if kind == _lib.OP_PHASE:
                                stage = pi.phase_scalar(mats, terms, base) * stage
                            else:
                                stage = pi.apply_u1_matrix(stage, tidx, perm[q0], pi.cond_matrix(mats, terms, base))
                            continue
                        stage = pi.apply_op(stage, tidx, kind, perm[q0], perm[q1] if kind != _lib.OP_U1 else 0, mats, mat)
                for off, slot in d["st"]:
                    gidx = base | off
                    r, l = gidx >> n_local, (gidx & lmask) | st_box
                    assert not written[r][l].any(), "two tiles of one sweep store the same amplitude"
                    written[r][l] = True
                    shards[r][l] = stage[slot:slot + len(st_box)]
                    peer_bytes += 16 * len(st_box) * (r != rank)
        for p in positions:
            live |= 1 << p
    return np.concatenate(shards), peer_bytes


@pytest.mark.parametrize("name,n,depth,world,onchip,tile", [
    ("syc", 12, 2, 2, 6, 7), ("syc", 12, 2, 4, 6, 7), ("qft", 11, 1, 2, 5, 6), ("hwe", 12, 2, 4, 6, 7), ("bv", 12, 1, 2, 6, 6),
    ("syc", 13, 3, 8, 6, 7),
])
def test_sharded_sweeps_move_peer_halves_through_the_tiles(name, n, depth, world, onchip, tile):
    """Several ranks, each owning the amplitudes whose top bits equal its rank: every tile has exactly one
    owner, sweeps whose tile holds rank bits read / write the peers' buffers, nothing unwritten is ever read
    and the concatenated shards equal the single-buffer result."""
    prog = _uncut_program(name, n, depth, onchip_max=onchip, stream_tile=tile)
    (plan,) = prog.plans()
    want = pi.run_plan(prog, plan, 0, return_state=True)
    got, peer_bytes = emulate_sharded(prog, plan, world)
    assert not np.isnan(got).any(), "the last sweep must leave every shard completely written"
    assert np.abs(got - want).max() < 1e-14
    n_local = n - (world.bit_length() - 1)
    touches = any(p >= n_local for pos, _, _ in plan.sweeps for p in pos)
    assert (peer_bytes > 0) == touches
