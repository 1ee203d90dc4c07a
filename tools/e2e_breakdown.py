"""Scratch: where does the end-to-end time of run_virtual_circuit_dense go? (host wall clock)"""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from importlib import import_module
PKG = "hardwareawareoptimalquantumcircuitcuttingandknitting_b200"
cutting = import_module(PKG + ".cutting"); vcm = import_module(PKG + ".virtual_circuit"); runm = import_module(PKG + ".run")
_lib = import_module(PKG + "._lib")
wl = sys.argv[1] if len(sys.argv) > 1 else "syc32d1"
dev = torch.device("cuda", 0)
circ, cut = cutting.make_baseline(wl, 0)
h = _lib.get_handle(0)
out = None
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    virt = vcm.VirtualCircuit(cut); t1 = time.perf_counter()
    frags = virt.active_fragments(); [virt.program(f).plans() for f in frags]; t2 = time.perf_counter()
    exs = [virt.executor(f, dev, True) for f in frags]; t3 = time.perf_counter()
    [e.upload() for e in exs]; t4 = time.perf_counter()
    tables = virt.simulate_fragments(dev); t5 = time.perf_counter()
    stats = torch.zeros(4, dtype=torch.float64, device=dev)
    if out is None:
        out = virt.knit_tables(tables, dev, stats=stats)
    else:
        virt.knit_tables(tables, dev, stats=stats, out=out)
    t6 = time.perf_counter()
    hs = stats.cpu(); t7 = time.perf_counter()
    print(f"{wl} it{it}: VirtualCircuit {1e3*(t1-t0):.2f} | compile {1e3*(t2-t1):.2f} | executors {1e3*(t3-t2):.2f} | upload {1e3*(t4-t3):.2f} | sim enqueue {1e3*(t5-t4):.2f} | knit enqueue {1e3*(t6-t5):.2f} | sync+d2h {1e3*(t7-t6):.2f} | total {1e3*(t7-t0):.2f} ms  launches {h.launch_count}")
