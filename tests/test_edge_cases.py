"""Edge cases of the path: ragged / degenerate inputs the reference's flow meets in practice.
Each case runs on the CPU through the plan interpreter and, with -m gpu, on the device."""
import math

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
vgm = import_module(f"{PKG}.virtual_gates")


def _qc(n, ncl=None):
    regs = [circuit.QuantumRegister(n, "q")]
    if ncl:
        regs.append(circuit.ClassicalRegister(ncl, "c"))
    return circuit.QuantumCircuit(*regs)


def case_unmeasured_clbits():
    """5 clbits, only 0, 2 and 4 are ever written: the dense result lives on the written bits."""
    qc = _qc(3, 5)
    qc.h(0); qc.cx(0, 1); qc.ry(0.7, 2); qc.cz(1, 2)
    qc.measure(0, 0); qc.measure(1, 2); qc.measure(2, 4)
    gi = [i for i, ins in enumerate(qc.data) if ins.operation.name == "cz"][0]
    return qc, cutting.apply_cuts(qc, cutting.CutSpec(gate_cuts=[gi]))


def case_fragment_without_measurement():
    """q2 is entangled with nothing that is measured and never measured itself (run.py:49-58)."""
    qc = _qc(3, 2)
    qc.h(0); qc.cx(0, 1); qc.h(2); qc.t(2)
    qc.measure(0, 0); qc.measure(1, 1)
    return qc, cutting.apply_cuts(qc, cutting.CutSpec())


def case_one_qubit_fragments():
    qc = _qc(2)
    qc.ry(0.9, 0); qc.rx(0.4, 1); qc.cx(0, 1); qc.h(0)
    qc.measure_all()
    return qc, cutting.apply_cuts(qc, cutting.CutSpec(gate_cuts=[2]))


def case_rzz_degenerate(theta):
    qc = _qc(2)
    qc.h(0); qc.ry(0.3, 1); qc.rzz(theta, 0, 1); qc.h(1)
    qc.measure_all()
    return qc, cutting.apply_cuts(qc, cutting.CutSpec(gate_cuts=[2]))


def case_vgate_inside_one_fragment():
    """A virtual gate whose two ends stay in the same fragment (qvm allows it)."""
    qc = _qc(3)
    qc.h(0); qc.cx(0, 1); qc.cz(1, 2); qc.cx(0, 2); qc.ry(0.2, 1)
    qc.measure_all()
    return qc, cutting.apply_cuts(qc, cutting.CutSpec(gate_cuts=[2], partitions=[[0, 1, 2]]))


def case_wire_cut_only_moved_qubit():
    """After the wire cut the new fragment holds nothing but the vmove qubit."""
    qc = _qc(2)
    qc.h(0); qc.cx(0, 1); qc.ry(0.6, 1); qc.rz(0.2, 1)
    qc.measure_all()
    return qc, cutting.apply_cuts(qc, cutting.CutSpec(wire_cuts=[(1, 1)]))


CASES = {
    "unmeasured_clbits": case_unmeasured_clbits,
    "fragment_without_measurement": case_fragment_without_measurement,
    "one_qubit_fragments": case_one_qubit_fragments,
    "rzz_theta_0": lambda: case_rzz_degenerate(0.0),
    "rzz_theta_pi": lambda: case_rzz_degenerate(math.pi),
    "rzz_theta_generic": lambda: case_rzz_degenerate(0.9),
    "vgate_inside_one_fragment": case_vgate_inside_one_fragment,
    "wire_cut_only_moved_qubit": case_wire_cut_only_moved_qubit,
}


def _expected(qc, cut):
    want, ov = oracle_knit(cut, 0.0)
    uncut = sv.exact_distribution(qc)
    assert max(abs(want.get(k, 0) - uncut.get(k, 0)) for k in set(want) | set(uncut)) < 1e-12
    return want


def _check_dict(got, want):
    keys = set(got) | set(want)
    assert max(abs(got.get(k, 0.0) - want.get(k, 0.0)) for k in keys) < 1e-10


@pytest.mark.parametrize("name", sorted(CASES))
def test_edge_case_host_compiler(name):
    qc, cut = CASES[name]()
    virt = vcm.VirtualCircuit(cut)
    want = _expected(qc, cut)
    tables = {f: pi.run_program(virt.program(f)) for f in virt.active_fragments()}
    for f in virt.active_fragments():       # instance de-duplication: copies equal their representatives' rows
        assert np.array_equal(tables[f], pi.run_program_deduped(virt.program(f)))
    masks, union = virt.output_masks()
    frags = list(tables)
    coeffs = [[c[0] for c in vg.knit_coefficients()] for vg in virt.vgates]
    dense = od.contract([tables[f] for f in frags], [virt._touches(f) for f in frags], coeffs,
                        [vcm._compress_mask(masks[f], union) for f in frags], bin(union).count("1"))
    keys = od.pdep(np.arange(len(dense), dtype=np.uint64), union)
    got = {int(k): float(v) for k, v in zip(keys.tolist(), dense.tolist()) if abs(v) > 1e-14}
    _check_dict(got, want)
    if name == "fragment_without_measurement":
        assert len(virt.active_fragments()) == len(virt.fragment_circuits) - 1
    if name in ("rzz_theta_0", "rzz_theta_pi"):
        assert virt.vgates[0].num_instantiations == 1 and virt.num_global_labels() == 1


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_edge_case_on_device(name):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    runm = import_module(f"{PKG}.run")
    qc, cut = CASES[name]()
    want = _expected(qc, cut)
    got, info = runm.run_virtual_circuit(vcm.VirtualCircuit(cut))
    _check_dict(got, want)
    # reference-faithful mode on the same inputs
    res5, _ = runm.run_virtual_circuit_dense(vcm.VirtualCircuit(cut), nearest=False, accuracy=1e-5)
    want5, _ = oracle_knit(cut, 1e-5)
    got5 = res5.to_dict()
    for k in set(got5) | set(want5):
        a, b = got5.get(k, 0.0), want5.get(k, 0.0)
        assert abs(a - b) < 1e-10 or abs(abs(a) - 1e-5) < 1e-9 or abs(abs(b) - 1e-5) < 1e-9
