"""Small end-to-end runs for compute-sanitizer (memcheck / racecheck): every kernel family once, at sizes a
sanitizer finishes in a minute.  `compute-sanitizer --tool memcheck python tests/sanitize_small.py`
(lives under tests/ because it checks against the oracle)."""
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import torch
from importlib import import_module

PKG = "hardwareawareoptimalquantumcircuitcuttingandknitting_b200"
cutting = import_module(PKG + ".cutting")
vcm = import_module(PKG + ".virtual_circuit")
runm = import_module(PKG + ".run")
gen = import_module(PKG + ".generators")
from conftest import make_semcheck_circuit
from oracle import statevector as sv

dev = torch.device("cuda", 0)
which = sys.argv[1:] or ["semcheck", "bv16", "add6", "syc16", "faithful"]


def check(name, circ, cut, **kw):
    dense, _ = runm.run_virtual_circuit_dense(vcm.VirtualCircuit(cut), device=dev, **kw)
    want = sv.dense(sv.exact_distribution(circ), circ.num_clbits)
    err = float(np.abs(dense.values.cpu().numpy() - want).max())
    print(f"{name}: max err {err:.2e}", flush=True)
    assert err < 1e-4 if kw.get("accuracy") else err < 1e-10, err


if "semcheck" in which:         # on-chip simulation, padded DMMA contraction, npd
    qc, cut = make_semcheck_circuit("cx")
    check("semcheck", qc, cut)
if "bv16" in which:             # wire cut, 9 | 8 qubits
    check("bv16", *cutting.make_baseline("bv16"))
if "add6" in which:             # two wire cuts on one qubit (recorded solver output, no z3 needed here)
    circ = gen.gen_circ("add", 6, 1).decompose_two_qubit()
    spec = cutting.CutSpec(wire_cuts=[(2, 10), (2, 53)], partitions=[[0, 1, 2], [3, 4, 5]])
    check("add6", circ, cutting.apply_cuts(circ, spec))
if "syc16" in which:            # K = 0: streaming (TMA) simulation of a 14-qubit fragment + outer-product knit
    c14 = gen.gen_circ("syc", 14, 2, seed=0).decompose_two_qubit()
    check("syc14-uncut", c14, cutting.apply_cuts(c14, cutting.CutSpec()), nearest=False)
if "faithful" in which:         # reference-faithful (pruned) knit
    qc, cut = make_semcheck_circuit("cz")
    check("semcheck-faithful", qc, cut, accuracy=1e-5)
print("sanitize_small ok")
