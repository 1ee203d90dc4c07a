"""Register-resident regime (one warp per instance, csrc/sim_warp_kernel.inc): which programs take it, and
the kernel's depth-first walk over measurement outcomes - restated in numpy, tests/plan_interpreter.py
run_plan_warp - against the same ops on the ancilla-enlarged state and against the oracle."""
import random

import numpy as np
import pytest

import plan_interpreter as pi
import test_random_circuits_cpu as rc
from conftest import make_semcheck_circuit
from hardwareawareoptimalquantumcircuitcuttingandknitting_b200 import _lib, circuit, compiler, cutting
from hardwareawareoptimalquantumcircuitcuttingandknitting_b200 import virtual_circuit as vcm
from oracle import dense as od
from oracle import instantiate as oi
from oracle import statevector as sv


def _check_program(prog, fold, n_labels=6, seed=0):
    rng = np.random.default_rng(seed)
    for plan in prog.plans(fold):
        assert plan.warp_base == prog.n_qubits and plan.n_state - plan.warp_base <= compiler.WARP_MAX_DEPTH
        assert set(plan.ops[:, 0].tolist()) <= {_lib.OP_U1, _lib.OP_CX, _lib.OP_CZ}
        for label in rng.choice(plan.labels, size=min(n_labels, len(plan.labels)), replace=False):
            want = pi.run_plan(prog, plan, int(label))
            got = pi.run_plan_warp(prog, plan, int(label))
            assert np.abs(got - want).max() < 1e-14


@pytest.mark.parametrize("cfg", ["bv16", "syc16d5", "hwe16d5"])
def test_baseline_fragments_take_the_register_regime(cfg):
    circ, cut = cutting.make_baseline(cfg, seed=1)
    virt = vcm.VirtualCircuit(cut)
    for f in virt.active_fragments():
        prog = virt.program(f)
        assert prog.warp and prog.n_qubits <= compiler.WARP_MAX_QUBITS
        _check_program(prog, True, n_labels=2)
    f = virt.active_fragments()[0]
    _check_program(virt.program(f), False, n_labels=1)          # unfolded rows: outcome bits select the column


@pytest.mark.parametrize("gname,theta", [("cx", None), ("cz", None), ("cy", None), ("rzz", 0.83), ("cp", 0.83)])
def test_semcheck_instances_vs_oracle(gname, theta):
    qc, cut = make_semcheck_circuit(gname, theta)
    virt = vcm.VirtualCircuit(cut)
    ov = oi.OracleVirtualCircuit(cut)
    K = len(virt.vgates)
    for f in virt.active_fragments():
        prog = virt.program(f)
        assert prog.warp
        labels = ov.instance_labels(f)
        for plan in prog.plans(True):
            for label in plan.labels:
                want = od.signed_fold(sv.exact_distribution(ov.instance(f, labels[int(label)])), ov.n_clbits, K,
                                      prog.out_mask)
                assert np.abs(pi.run_plan_warp(prog, plan, int(label)) - want).max() < 1e-12


def test_mid_circuit_measurement_of_the_input_is_a_branch_point():
    qc = circuit.QuantumCircuit(circuit.QuantumRegister(2, "q"), circuit.ClassicalRegister(3, "c"))
    qc.h(0); qc.measure(0, 2); qc.h(0); qc.cx(0, 1); qc.ry(0.3, 1); qc.measure(0, 0); qc.measure(1, 1)
    virt = vcm.VirtualCircuit(qc)
    (f,) = virt.active_fragments()
    prog = virt.program(f)
    assert prog.warp
    (plan,) = prog.plans(True)
    assert plan.warp_base == 2 and plan.n_state == 3
    want = sv.dense(sv.exact_distribution(qc), 3)
    assert np.abs(pi.run_plan_warp(prog, plan, 0) - want).max() < 1e-14


def test_programs_that_stay_on_the_other_kernels():
    # a general two-qubit unitary (here: rzz) cannot run in the register regime
    qc = circuit.QuantumCircuit(circuit.QuantumRegister(3, "q"))
    qc.h(0); qc.rzz(0.4, 0, 1); qc.cx(1, 2)
    qc.measure_all()
    virt = vcm.VirtualCircuit(qc)
    (f,) = virt.active_fragments()
    assert not virt.program(f).warp
    # 11 qubits: too wide
    wide = circuit.QuantumCircuit(circuit.QuantumRegister(11, "q"))
    for q in range(10):
        wide.cx(q, q + 1)
    wide.measure_all()
    v2 = vcm.VirtualCircuit(wide)
    assert not v2.program(v2.active_fragments()[0]).warp
    # knob
    assert not compiler.FragmentProgram(virt.fragment_circuits[f], f, qc.num_clbits, warp=False).warp


@pytest.mark.parametrize("seed", range(6))
def test_random_cut_circuits(seed):
    rng = random.Random(900 + seed)
    n = rng.randint(4, 9)
    qc = rc.random_circuit(rng, n, rng.randint(20, 40))
    cut = cutting.apply_cuts(qc, rc.random_cut(rng, qc, max_gate_cuts=2, wire_cut=(seed % 2 == 0)))
    virt = vcm.VirtualCircuit(cut)
    for f in virt.active_fragments():
        prog = virt.program(f)
        if prog.warp:          # (fragments holding a general two-qubit gate keep the other kernels)
            _check_program(prog, True, n_labels=3, seed=seed)


# ------------------------------------------------------------------ tree-walk simulation (sim_tree_kernel.inc)
@pytest.mark.parametrize("gname,theta", [("cx", None), ("cz", None), ("cy", None), ("rzz", 0.83), ("cp", 0.83)])
def test_tree_tables_equal_per_instance_tables(gname, theta):
    """The level-by-level tree walk (every shared prefix once, one representative per class of identical
    variants, partial rows per outcome leaf, per-label combine) gives the table the per-instance programs give."""
    qc, cut = make_semcheck_circuit(gname, theta)
    virt = vcm.VirtualCircuit(cut)
    for f in virt.active_fragments():
        prog = virt.program(f)
        tree = prog.tree()
        assert tree is not None
        want = pi.run_program(prog, True)
        got = pi.run_tree(prog)
        assert np.abs(got - want).max() < 1e-14


def test_tree_on_a_baseline_fragment_and_with_mid_circuit_measurement():
    circ, cut = cutting.make_baseline("bv16", seed=0)
    virt = vcm.VirtualCircuit(cut)
    for f in virt.active_fragments():
        prog = virt.program(f)
        assert prog.tree() is not None
        assert np.abs(pi.run_tree(prog) - pi.run_program(prog, True)).max() < 1e-14
    # mid-circuit measurement of the input circuit next to a gate cut: a level whose outcome is a column bit
    q4 = circuit.QuantumCircuit(circuit.QuantumRegister(4, "q"), circuit.ClassicalRegister(5, "c"))
    for q in range(4):
        q4.ry(0.3 + 0.2 * q, q)
    q4.cx(0, 1); q4.measure(0, 4); q4.h(0); q4.cx(2, 3); q4.cx(1, 2); q4.rx(0.4, 1); q4.cx(0, 1); q4.ry(0.2, 2)
    for q in range(4):
        q4.measure(q, q)
    gidx = [i for i, ins in enumerate(q4.data) if ins.operation.name == "cx"][2]
    virt = vcm.VirtualCircuit(cutting.apply_cuts(q4, cutting.CutSpec(gate_cuts=[gidx])))
    kinds = set()
    for f in virt.active_fragments():
        prog = virt.program(f)
        tree = prog.tree()
        assert tree is not None
        kinds |= {l.kind for l in tree.levels}
        assert np.abs(pi.run_tree(prog) - pi.run_program(prog, True)).max() < 1e-14
    assert _lib.TREE_MMEAS in kinds and _lib.TREE_SLOT in kinds


@pytest.mark.parametrize("seed", range(6))
def test_tree_random_cut_circuits(seed):
    rng = random.Random(1300 + seed)
    n = rng.randint(4, 8)
    qc = rc.random_circuit(rng, n, rng.randint(15, 30))
    cut = cutting.apply_cuts(qc, rc.random_cut(rng, qc, max_gate_cuts=2, wire_cut=(seed % 2 == 0)))
    virt = vcm.VirtualCircuit(cut)
    for f in virt.active_fragments():
        prog = virt.program(f)
        if prog.tree() is not None and prog.num_labels <= 64:
            assert np.abs(pi.run_tree(prog) - pi.run_program(prog, True)).max() < 1e-13
            lo, hi = prog.num_labels // 3, prog.num_labels
            part = pi.run_tree(prog, (lo, hi))
            assert np.abs(part[lo:hi] - pi.run_program(prog, True)[lo:hi]).max() < 1e-13 and not part[:lo].any()
