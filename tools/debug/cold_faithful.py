"""Scratch: where does the cold end-to-end call of the reference-default mode spend its time?"""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from importlib import import_module
PKG = "hardwareawareoptimalquantumcircuitcuttingandknitting_b200"
cutting = import_module(PKG + ".cutting"); vcm = import_module(PKG + ".virtual_circuit"); runm = import_module(PKG + ".run")
_lib = import_module(PKG + "._lib")
wl = sys.argv[1] if len(sys.argv) > 1 else "syc16d5"
dev = torch.device("cuda", 0)
circ, cut = cutting.make_baseline(wl, 0)
h = _lib.get_handle(0)
for it in range(6):
    cold = it % 2 == 0
    if cold:
        vcm.clear_program_cache()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    virt = vcm.VirtualCircuit(cut); t1 = time.perf_counter()
    frags = virt.active_fragments(); t2 = time.perf_counter()
    exs = [virt.executor(f, dev, False) for f in frags]; t3 = time.perf_counter()
    tables = virt.simulate_fragments(dev, fold=False); t4 = time.perf_counter()
    torch.cuda.synchronize(); t5 = time.perf_counter()
    out = virt.knit_tables_faithful(tables, 1e-5, dev); t6 = time.perf_counter()
    torch.cuda.synchronize(); t7 = time.perf_counter()
    print(f"{wl} {'cold' if cold else 'warm'}: VirtualCircuit {1e3*(t1-t0):.2f} | programs {1e3*(t2-t1):.2f} | executors {1e3*(t3-t2):.2f} | sim enqueue {1e3*(t4-t3):.2f} | sim wait {1e3*(t5-t4):.2f} | knit enqueue {1e3*(t6-t5):.2f} | knit wait {1e3*(t7-t6):.2f} | total {1e3*(t7-t0):.2f} ms")
