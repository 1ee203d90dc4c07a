"""Host compiler (templates, slots, patterns, ancillas, sweeps, fold masks) against the oracle,
executed by the numpy plan interpreter (tests/plan_interpreter.py).  CPU only."""
import numpy as np
import pytest

import plan_interpreter as pi
from conftest import make_semcheck_circuit
from oracle import dense as od
from oracle import instantiate as oi
from oracle import statevector as sv

PKG = "hardwareawareoptimalquantumcircuitcuttingandknitting_b200"
from importlib import import_module

cutting = import_module(f"{PKG}.cutting")
vcm = import_module(f"{PKG}.virtual_circuit")
compiler = import_module(f"{PKG}.compiler")
circuit = import_module(f"{PKG}.circuit")


def _tables(virt):
    return {f: pi.run_program(virt.program(f)) for f in virt.active_fragments()}


def _contract(virt, tables):
    masks, union = virt.output_masks()
    frags = list(tables)
    coeffs = [[c[0] for c in vg.knit_coefficients()] for vg in virt.vgates]
    return od.contract([tables[f] for f in frags], [virt._touches(f) for f in frags], coeffs,
                       [vcm._compress_mask(masks[f], union) for f in frags], bin(union).count("1"))


@pytest.mark.parametrize("gname,theta", [("cx", None), ("cz", None), ("cy", None), ("rzz", 0.83), ("cp", 0.83)])
def test_semcheck_every_instance_and_knit(gname, theta):
    qc, cut = make_semcheck_circuit(gname, theta)
    virt = vcm.VirtualCircuit(cut)
    ov = oi.OracleVirtualCircuit(cut)
    tables = _tables(virt)
    K = len(virt.vgates)
    for f in virt.active_fragments():
        prog = virt.program(f)
        labels = ov.instance_labels(f)
        assert labels == virt.get_instance_labels(f)            # bit-exact enumeration
        for li, lab in enumerate(labels):
            want = od.signed_fold(sv.exact_distribution(ov.instance(f, lab)), ov.n_clbits, K, prog.out_mask)
            assert np.abs(want - tables[f][li]).max() < 1e-14
    res = _contract(virt, tables)
    uncut = sv.dense(sv.exact_distribution(qc), qc.num_clbits)
    if gname != "cp":
        assert np.abs(res - uncut).max() < 1e-13


def test_unfolded_rows_carry_config_bits():
    qc, cut = make_semcheck_circuit("cx")
    virt = vcm.VirtualCircuit(cut)
    ov = oi.OracleVirtualCircuit(cut)
    K = len(virt.vgates)
    for f in virt.active_fragments():
        prog = virt.program(f)
        table = pi.run_program(prog, fold=False)
        m = bin(prog.out_mask).count("1")
        for li, lab in enumerate(ov.instance_labels(f)):
            dist = sv.exact_distribution(ov.instance(f, lab))
            want = np.zeros(1 << (m + len(prog.radix)))
            for key, v in dist.items():
                x = int(od.pext(np.uint64(key & ((1 << ov.n_clbits) - 1)), prog.out_mask))
                cfg = key >> ov.n_clbits
                idx = x
                for d, k in enumerate(prog.vgate_indices):
                    idx |= ((cfg >> k) & 1) << (m + d)
                want[idx] += v
            assert np.abs(want - table[li]).max() < 1e-14


def test_bv16_wire_cut():
    circ, cut = cutting.make_baseline("bv16")
    virt = vcm.VirtualCircuit(cut)
    assert sorted(len(f) for f in virt.fragment_circuits) == [8, 9]
    assert virt.num_global_labels() == 8
    res = _contract(virt, _tables(virt))
    want = np.zeros(1 << 16)
    want[(1 << 16) - 1] = 1.0                      # secret = fifteen ones, ancilla measured as 1
    assert np.abs(res - want).max() < 1e-13


def test_streaming_schedule_equals_onchip():
    """The same program forced into sweeps (tile 6) gives the same rows."""
    circ, cut = cutting.make_baseline("bv16")
    virt = vcm.VirtualCircuit(cut)
    for f in virt.active_fragments():
        a = compiler.FragmentProgram(virt.fragment_circuits[f], f, virt.num_clbits)
        b = compiler.FragmentProgram(virt.fragment_circuits[f], f, virt.num_clbits, onchip_max=4, stream_tile=7)
        assert any(len(p.sweeps) > 1 for p in b.plans())
        for p in b.plans():
            for positions, _, _ in p.sweeps:
                assert positions[:5] == [0, 1, 2, 3, 4] and positions == sorted(positions)
        assert np.abs(pi.run_program(a) - pi.run_program(b)).max() < 1e-14


def test_mid_circuit_measurement_of_input_circuit():
    qc = circuit.QuantumCircuit(circuit.QuantumRegister(2, "q"), circuit.ClassicalRegister(3, "c"))
    qc.h(0); qc.cx(0, 1); qc.measure(0, 0); qc.h(0); qc.ry(0.4, 1); qc.measure(0, 1); qc.measure(1, 2)
    virt = vcm.VirtualCircuit(qc)
    (f,) = virt.active_fragments()
    row = pi.run_program(virt.program(f))[0]
    want = sv.dense(sv.exact_distribution(qc), 3)
    assert np.abs(row - want).max() < 1e-14
    # the masks the knit and DenseResult.key_mask use cover EVERY written clbit, the mid-circuit one included
    prog = virt.program(f)
    assert prog.out_mask == 0b111 and prog.out_clbits == [0, 1, 2] and prog.row_len() == 8
    masks, union = virt.output_masks()
    assert union == 0b111
    # mid-circuit outcome on the HIGHEST clbit, terminal ones below it
    qc2 = circuit.QuantumCircuit(circuit.QuantumRegister(2, "q"), circuit.ClassicalRegister(3, "c"))
    qc2.h(0); qc2.measure(0, 2); qc2.h(0); qc2.cx(0, 1); qc2.measure(0, 0); qc2.measure(1, 1)
    v2 = vcm.VirtualCircuit(qc2)
    (f2,) = v2.active_fragments()
    p2 = v2.program(f2)
    assert p2.out_mask == 0b111 and p2.row_len() == 8
    assert np.abs(pi.run_program(p2)[0] - sv.dense(sv.exact_distribution(qc2), 3)).max() < 1e-14


def test_errors_mirror_reference():
    qc = circuit.QuantumCircuit(circuit.QuantumRegister(1, "a"), circuit.QuantumRegister(1, "b"))
    qc.cx(qc.qregs[0][0], qc.qregs[1][0])
    with pytest.raises(ValueError, match="multiple fragments"):
        vcm.VirtualCircuit(qc)
    ok = circuit.QuantumCircuit(circuit.QuantumRegister(1, "a"))
    v = vcm.VirtualCircuit(ok)
    with pytest.raises(ValueError, match="Fragment not found"):
        v.get_backend(circuit.QuantumRegister(1, "zz"))
    with pytest.raises(ValueError, match="Fragment not found"):
        v.set_backend(circuit.QuantumRegister(1, "zz"), object())


def test_baseline_config_shapes():
    shapes = {"bv16": ([8, 9], 1, 8), "hwe16d5": ([8, 8], 5, 7776), "syc16d5": ([8, 8], 4, 1296),
              "syc32d1": ([14, 18], 0, 1)}
    for cfg, (frag_sizes, K, L) in shapes.items():
        _, cut = cutting.make_baseline(cfg)
        v = vcm.VirtualCircuit(cut)
        assert sorted(len(f) for f in v.fragment_circuits) == frag_sizes
        assert len(v.vgates) == K and v.num_global_labels() == L
    _, cut = cutting.make_baseline("syc32d1")
    v = vcm.VirtualCircuit(cut)
    masks, union = v.output_masks()
    assert sorted(masks.values()) == [0x7EFF0000, 0x8100FFFF] and union == 0xFFFFFFFF


def test_c_cluster_scheduler_equals_python_reference():
    """qck_host_cluster_ops (C) against the Python original on every on-chip plan of the baseline configs
    and on random op lists (n_live growth, lone two-qubit ops, clusters of every size)."""
    import random
    import cluster_reference as cr
    _lib = import_module(f"{PKG}._lib")
    cases = []
    for cfg in ("bv16", "syc16d5", "hwe16d5"):
        circ, cut = cutting.make_baseline(cfg)
        virt = vcm.VirtualCircuit(cut)
        for f in virt.active_fragments():
            prog = compiler.FragmentProgram(virt.fragment_circuits[f], f, cut.num_clbits, cluster=False)
            for plan in prog.plans()[::7]:
                cases.append((plan.ops, plan.sweeps))
    rng = random.Random(11)
    for _ in range(60):
        T = rng.randint(1, 9)
        rows = []
        for _ in range(rng.randint(1, 70)):
            kind = rng.choice([_lib.OP_U1, _lib.OP_CX, _lib.OP_CZ, _lib.OP_U2]) if T >= 2 else _lib.OP_U1
            q = rng.sample(range(T), 2) if kind != _lib.OP_U1 else [rng.randrange(T), 0]
            rows.append([kind, q[0], q[1], rng.randrange(100) * 8, -1, 0, rng.choice([0, 0, 2, 3, 4, T]), 0])
        cases.append((np.asarray(rows, dtype=np.int32), [(list(range(T)), 0, len(rows))]))
    for ops, sweeps in cases:
        got_ops, got_sw = compiler._cluster_sweeps(ops, sweeps)
        want_ops, want_sw = cr.cluster_sweeps_reference(ops, sweeps)
        assert got_sw == want_sw
        assert np.array_equal(got_ops, want_ops)


@pytest.mark.parametrize("gname,theta", [("cx", None), ("rzz", 0.83)])
def test_shared_prefix_split_does_not_change_rows(gname, theta):
    """On-chip programs split into a label-independent prefix (run once) and the per-instance rest
    (compiler.SHARE_PREFIX, qck.h QCK_SWEEP_SHARED): same rows as the unsplit program, and the prefix holds no
    label-dependent op (asserted by the interpreter)."""
    qc, cut = make_semcheck_circuit(gname, theta)
    virt = vcm.VirtualCircuit(cut)
    n_split = 0
    for f in virt.active_fragments():
        circ = virt.fragment_circuits[f]
        plain = compiler.FragmentProgram(circ, f, virt.num_clbits, share_prefix=False)
        split = compiler.FragmentProgram(circ, f, virt.num_clbits, share_prefix=True)
        for fold in (True, False):
            assert all(len(p.sweeps) == 1 and not p.shared_prefix for p in plain.plans(fold))
            n_split += sum(p.shared_prefix for p in split.plans(fold))
            for p in split.plans(fold):
                if p.shared_prefix:
                    (pos0, b0, e0), (pos1, b1, e1) = p.sweeps
                    assert b0 == 0 and e0 == b1 and e1 == len(p.ops) and e0 > 0
            assert np.abs(pi.run_program(plain, fold=fold) - pi.run_program(split, fold=fold)).max() < 1e-14
    assert n_split > 0


def test_shared_prefix_auto_policy():
    # auto: only with many instances and a prefix that is most of the program - never for the semcheck circuit
    qc, cut = make_semcheck_circuit("cx")
    virt = vcm.VirtualCircuit(cut)
    for f in virt.active_fragments():
        prog = compiler.FragmentProgram(virt.fragment_circuits[f], f, virt.num_clbits, share_prefix="auto")
        assert not any(p.shared_prefix for p in prog.plans())


def test_identical_instances_are_found_and_bit_identical():
    """Several instantiations of a virtual gate look the same from one side (the I and the Z term of a wire cut
    both measure Z; virtual_gates.py:62-103): bv-16's sender side needs 4 of its 8 instances, the receiver 6."""
    circ, cut = cutting.make_baseline("bv16")
    virt = vcm.VirtualCircuit(cut)
    uniq = sorted(len(np.unique(virt.program(f).canonical_labels())) for f in virt.active_fragments())
    assert uniq == [4, 6]
    for f in virt.active_fragments():
        prog = virt.program(f)
        assert np.array_equal(pi.run_program(prog), pi.run_program_deduped(prog))
    # gate cuts: 5 of the 6 instantiations of a cut cx differ on either side
    circ, cut = cutting.make_baseline("syc16d5")
    virt = vcm.VirtualCircuit(cut)
    for f in virt.active_fragments():
        prog = virt.program(f)
        assert len(np.unique(prog.canonical_labels())) == 5 ** 4 and prog.num_labels == 6 ** 4


def _all_cut_circuits():
    import random
    import test_edge_cases as tec
    import test_random_circuits_cpu as trc
    for gname, theta in [("cx", None), ("cz", None), ("cy", None), ("rzz", 0.83), ("cp", 0.83)]:
        yield make_semcheck_circuit(gname, theta)[1]
    for cfg in ("bv16", "syc16d5", "hwe16d5"):
        yield cutting.make_baseline(cfg)[1]
    for name in sorted(tec.CASES):
        yield tec.CASES[name]()[1]
    for seed in range(12):
        rng = random.Random(1000 + seed)
        qc = trc.random_circuit(rng, rng.randint(4, 6), rng.randint(10, 22))
        yield cutting.apply_cuts(qc, trc.random_cut(rng, qc, max_gate_cuts=2, wire_cut=(seed % 3 == 0)))


def test_template_build_plan_equals_the_emit_loop():
    """compiler._build_plan (rows built once per program, selected and renumbered per pattern) against the original
    per-pattern emit loop (tests/build_plan_reference.py): identical op arrays, masks and output positions."""
    import build_plan_reference as ref
    n_plans = 0
    for cut in _all_cut_circuits():
        virt = vcm.VirtualCircuit(cut)
        for f in virt.fragment_circuits:
            prog = compiler.FragmentProgram(virt.fragment_circuits[f], f, virt.num_clbits, cluster=False,
                                            share_prefix=False)
            for fold in (True, False):
                for plan in prog.plans(fold):
                    if plan.n_state > prog.onchip_max:
                        continue
                    ops, n_state, out_pos, sum_mask, sign_mask = ref.emit_ops(prog, plan.pattern, fold)
                    assert np.array_equal(plan.ops, ops)
                    assert (plan.n_state, plan.out_pos, plan.sum_mask, plan.sign_mask) == \
                        (n_state, out_pos, sum_mask, sign_mask)
                    n_plans += 1
    assert n_plans > 200
