"""Scratch: cProfile of the warm end-to-end call."""
import sys, os, cProfile, pstats, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from importlib import import_module
PKG = "hardwareawareoptimalquantumcircuitcuttingandknitting_b200"
cutting = import_module(PKG + ".cutting"); vcm = import_module(PKG + ".virtual_circuit"); runm = import_module(PKG + ".run")
wl = sys.argv[1] if len(sys.argv) > 1 else "syc32d1"
dev = torch.device("cuda", 0)
circ, cut = cutting.make_baseline(wl, 0)
out = None
for _ in range(3):
    r, _ = runm.run_virtual_circuit_dense(vcm.VirtualCircuit(cut), device=dev, nearest=True, out=out)
    out = r.values
virts = [vcm.VirtualCircuit(cut) for _ in range(20)]
torch.cuda.synchronize()
t0 = time.perf_counter()
for v in [vcm.VirtualCircuit(cut) for _ in range(20)]:
    runm.run_virtual_circuit_dense(v, device=dev, nearest=True, out=out)
torch.cuda.synchronize()
print(f"{wl}: {(time.perf_counter()-t0)/20*1e3:.3f} ms per call (plain, incl. VirtualCircuit construction)")
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
for v in virts:
    runm.run_virtual_circuit_dense(v, device=dev, nearest=True, out=out)
pr.disable()
print(f"{wl}: {(time.perf_counter()-t0)/20*1e3:.3f} ms per call (under cProfile)")
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
