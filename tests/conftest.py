import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def golden():
    return load_golden


def make_semcheck_circuit(gname, theta=None):
    """The 4-qubit circuit of tests/golden/make_golden.py::semcheck (one wire cut + one gate cut)."""
    from hardwareawareoptimalquantumcircuitcuttingandknitting_b200.circuit import QuantumCircuit, QuantumRegister
    from hardwareawareoptimalquantumcircuitcuttingandknitting_b200.cutting import CutSpec, apply_cuts
    qc = QuantumCircuit(QuantumRegister(4, "q"))
    for q in range(4):
        qc.ry(0.3 + 0.4 * q, q)
        qc.rz(0.2 * q + 0.1, q)
    qc.cx(0, 1); qc.h(1); qc.cx(1, 2); qc.rx(0.7, 1); qc.cx(2, 3)
    if theta is None:
        getattr(qc, gname)(0, 3)
    else:
        getattr(qc, gname)(theta, 0, 3)
    qc.ry(0.5, 0); qc.rx(0.4, 3); qc.h(2)
    qc.measure_all()
    gidx = [i for i, ins in enumerate(qc.data) if ins.operation.name == gname][-1]
    return qc, apply_cuts(qc, CutSpec(gate_cuts=[gidx], wire_cuts=[(1, 8)]))


def oracle_knit(cut, acc=0.0):
    """Per-instance exact distributions (oracle simulator) + reference-order sparse knit."""
    from oracle import instantiate as oi, qpd_tables as qt, sparse_knit as sk, statevector as sv
    ov = oi.OracleVirtualCircuit(cut)
    res, touches = [], []
    for frag in ov.fragments:
        labels = ov.instance_labels(frag)
        if not any(ov.has_measurement(frag, l) for l in labels):
            continue
        res.append([sk.prune(sv.exact_distribution(ov.instance(frag, l)), acc) for l in labels])
        touches.append(ov.touches(frag))
    vg = [(k, qt.knit_param(k, th), r) for (k, th, _), r in zip(ov.vgates, ov.radices)]
    return sk.knit(res, touches, vg, ov.n_clbits, acc), ov
