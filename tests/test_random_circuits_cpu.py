"""Randomised cut circuits (all virtual-gate kinds, wire cuts, 2-4 fragments): the compiled device
programs (numpy plan interpreter) + closed-form contraction against the oracle's instance-by-instance
simulation + sparse reference-order knit.  CPU only; seeds are fixed."""
import random

import numpy as np
import pytest

import plan_interpreter as pi
from conftest import oracle_knit
from oracle import dense as od
from oracle import statevector as sv

PKG = "hardwareawareoptimalquantumcircuitcuttingandknitting_b200"
from importlib import import_module

circuit = import_module(f"{PKG}.circuit")
cutting = import_module(f"{PKG}.cutting")
vcm = import_module(f"{PKG}.virtual_circuit")

ONE_Q = ["h", "x", "y", "z", "s", "sdg", "t", "tdg", "sx"]
ONE_Q_PARAM = ["rx", "ry", "rz", "p"]
TWO_Q = ["cx", "cz", "cy", "swap"]
TWO_Q_PARAM = ["cp", "rzz"]
CUTTABLE = ("cx", "cz", "cy", "cp", "rzz")


def random_circuit(rng, n, depth):
    qc = circuit.QuantumCircuit(circuit.QuantumRegister(n, "q"))
    for _ in range(depth):
        r = rng.random()
        if r < 0.35:
            getattr(qc, rng.choice(ONE_Q))(rng.randrange(n))
        elif r < 0.55:
            getattr(qc, rng.choice(ONE_Q_PARAM))(rng.uniform(-3, 3), rng.randrange(n))
        elif r < 0.6:
            qc.u(rng.uniform(-3, 3), rng.uniform(-3, 3), rng.uniform(-3, 3), rng.randrange(n))
        else:
            a, b = rng.sample(range(n), 2)
            if rng.random() < 0.7:
                getattr(qc, rng.choice(TWO_Q))(a, b)
            else:
                getattr(qc, rng.choice(TWO_Q_PARAM))(rng.uniform(-3, 3), a, b)
    qc.measure_all()
    return qc


def random_cut(rng, qc, max_gate_cuts, wire_cut):
    two = [i for i, ins in enumerate(qc.data) if ins.operation.name in CUTTABLE]
    gate_cuts = sorted(rng.sample(two, min(len(two), rng.randint(1, max_gate_cuts)))) if two else []
    wire_cuts = []
    if wire_cut:
        cands = [i for i, ins in enumerate(qc.data)
                 if ins.operation.name not in ("measure", "barrier") and i not in gate_cuts]
        if cands:
            i = rng.choice(cands)
            q = qc.qubit_index(rng.choice(qc.data[i].qubits))
            wire_cuts.append((q, i))
    return cutting.CutSpec(gate_cuts=gate_cuts, wire_cuts=wire_cuts)


def product_dense(virt):
    tables = {f: pi.run_program(virt.program(f)) for f in virt.active_fragments()}
    masks, union = virt.output_masks()
    frags = list(tables)
    coeffs = [[c[0] for c in vg.knit_coefficients()] for vg in virt.vgates]
    return od.contract([tables[f] for f in frags], [virt._touches(f) for f in frags], coeffs,
                       [vcm._compress_mask(masks[f], union) for f in frags], bin(union).count("1")), union


@pytest.mark.parametrize("seed", range(12))
def test_random_cut_circuit(seed):
    rng = random.Random(1000 + seed)
    n = rng.randint(4, 6)
    qc = random_circuit(rng, n, rng.randint(10, 22))
    spec = random_cut(rng, qc, max_gate_cuts=2, wire_cut=(seed % 3 == 0))
    cut = cutting.apply_cuts(qc, spec)
    virt = vcm.VirtualCircuit(cut)
    got, union = product_dense(virt)
    want_d, ov = oracle_knit(cut, 0.0)
    assert union == (1 << n) - 1
    want = np.zeros(1 << n)
    for k, v in want_d.items():
        want[k] = v
    assert np.abs(got - want).max() < 1e-12, (seed, [len(r) for r in cut.qregs], len(virt.vgates))
    # labels bit-exact per fragment
    for f in virt.fragment_circuits:
        assert virt.get_instance_labels(f) == ov.instance_labels(f)
    # instance de-duplication: a row equals the row of its representative BIT FOR BIT, and representatives are
    # fixed points of the map
    for f in virt.active_fragments():
        prog = virt.program(f)
        src = prog.canonical_labels()
        assert np.array_equal(src[src], src) and np.all(src <= np.arange(len(src)))
        full = pi.run_program(prog)
        assert np.array_equal(full, pi.run_program_deduped(prog))
        assert np.array_equal(pi.run_program(prog, fold=False), pi.run_program_deduped(prog, fold=False))
    has_cp = any(type(v).__name__ == "VirtualCPhase" for v in virt.vgates)
    if not has_cp:      # every decomposition except the reference's CPhase reproduces the uncut circuit
        uncut = sv.dense(sv.exact_distribution(qc), n)
        assert np.abs(got - uncut).max() < 1e-12


def test_fusion_and_clustering_do_not_change_rows():
    compiler = import_module(f"{PKG}.compiler")
    rng = random.Random(7)
    for trial in range(6):
        qc = random_circuit(rng, 5, 25)
        virt = vcm.VirtualCircuit(qc)
        (f,) = virt.active_fragments()
        circ = virt.fragment_circuits[f]
        base = pi.run_program(compiler.FragmentProgram(circ, f, virt.num_clbits, cluster=False, fuse=False))
        for kw in ({"cluster": True, "fuse": False}, {"cluster": False, "fuse": True}, {"cluster": True, "fuse": True},
                   {"cluster": True, "fuse": True, "onchip_max": 3, "stream_tile": 4}):
            got = pi.run_program(compiler.FragmentProgram(circ, f, virt.num_clbits, **kw))
            assert np.abs(got - base).max() < 1e-13, kw
        want = sv.dense(sv.exact_distribution(qc), 5)
        assert np.abs(base[0] - want).max() < 1e-13


def test_program_cache_keyed_by_structure():
    vcm.clear_program_cache()
    a = circuit.QuantumCircuit(circuit.QuantumRegister(2, "q")); a.h(0); a.cx(0, 1); a.rz(0.3, 1); a.measure_all()
    b = circuit.QuantumCircuit(circuit.QuantumRegister(2, "q")); b.h(0); b.cx(0, 1); b.rz(0.3, 1); b.measure_all()
    c = circuit.QuantumCircuit(circuit.QuantumRegister(2, "q")); c.h(0); c.cx(0, 1); c.rz(0.31, 1); c.measure_all()
    va, vb, vc_ = vcm.VirtualCircuit(a), vcm.VirtualCircuit(b), vcm.VirtualCircuit(c)
    pa = va.program(va.active_fragments()[0])
    pb = vb.program(vb.active_fragments()[0])
    pc = vc_.program(vc_.active_fragments()[0])
    assert pa is pb and pa is not pc
    vcm.clear_program_cache()
    assert vcm.VirtualCircuit(a).program(vcm.VirtualCircuit(a).active_fragments()[0]) is not pa
