"""Scratch: uncut syc-N d1 statevector run for profiling the streaming kernels."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from importlib import import_module
PKG = "hardwareawareoptimalquantumcircuitcuttingandknitting_b200"
gen = import_module(PKG + ".generators"); vcm = import_module(PKG + ".virtual_circuit"); _lib = import_module(PKG + "._lib")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 28
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
circ = gen.gen_circ("syc", n, 1, seed=0).decompose_two_qubit()
virt = vcm.VirtualCircuit(circ)
(frag,) = virt.active_fragments()
dev = torch.device("cuda", 0)
ex = virt.executor(frag, dev, True)
h = _lib.get_handle(0)
t = ex.run(h)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    ex.run(h, out=t)
e1.record(); torch.cuda.synchronize()
pl = ex.plans[0]
print(f"syc-{n} uncut: {e0.elapsed_time(e1)/reps:.2f} ms, sweeps {len(pl.sweeps)}, records {len(pl.ops)}, sum {t.sum().item():.12f}")
for pos, b, e in pl.sweeps:
    print("  sweep tile", pos, "records", e - b)
