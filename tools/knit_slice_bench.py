"""Scratch: knit_outer on 1, 1/2, 1/4, 1/8 of the syc-32 d1 output on ONE GPU (what a rank of a 2 / 4 / 8-GPU run
writes, without the peer exchange): fixed cost per launch against streaming rate."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
from importlib import import_module
PKG = "hardwareawareoptimalquantumcircuitcuttingandknitting_b200"
cutting = import_module(PKG + ".cutting"); vcm = import_module(PKG + ".virtual_circuit")
dev = torch.device("cuda", 0)
circ, cut = cutting.make_baseline("syc32d1", 0)
virt = vcm.VirtualCircuit(cut)
tables = virt.simulate_fragments(dev)
out = torch.empty(1 << 32, dtype=torch.float64, device=dev)
stats = torch.zeros(4, dtype=torch.float64, device=dev)
for frac in (32, 8, 4, 1):
    n = (1 << 32) // frac
    for which in (0, frac - 1):
        y0 = which * n
        for it in range(3):
            virt.knit_tables(tables, dev, stats=stats, y_range=(y0, y0 + n), out=out[:n])
        torch.cuda.synchronize()
        reps = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for it in range(reps):   # back to back: the host runs ahead, the GPU never waits for an enqueue
            virt.knit_tables(tables, dev, stats=stats, y_range=(y0, y0 + n), out=out[:n])
        e1.record(); e1.synchronize()
        t = e0.elapsed_time(e1) / reps
        print(f"1/{frac} slice #{which}: {t:.4f} ms per launch back to back = {8 * n / t / 1e6:.0f} GB/s")
