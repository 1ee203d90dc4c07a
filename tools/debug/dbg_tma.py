"""Scratch: isolate a TMA-sweep problem on small tiles (prints progress per fragment)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import numpy as np, torch
from importlib import import_module
PKG = "hardwareawareoptimalquantumcircuitcuttingandknitting_b200"
cutting = import_module(f"{PKG}.cutting"); vcm = import_module(f"{PKG}.virtual_circuit")
compiler = import_module(f"{PKG}.compiler"); _lib = import_module(f"{PKG}._lib")
cfg, onchip, tile = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda", 0)
circ, cut = cutting.make_baseline(cfg)
virt = vcm.VirtualCircuit(cut)
h = _lib.get_handle(0)
for f in virt.active_fragments():
    a = compiler.FragmentProgram(virt.fragment_circuits[f], f, virt.num_clbits)
    b = compiler.FragmentProgram(virt.fragment_circuits[f], f, virt.num_clbits, onchip_max=onchip, stream_tile=tile)
    ra = compiler.FragmentExecutor(a, dev).run(h).cpu().numpy()
    eb = compiler.FragmentExecutor(b, dev)
    for p in eb.plans:
        print("plan n_state", p.n_state, "labels", len(p.labels), "sweeps", [(pos, e - bb) for pos, bb, e in p.sweeps], flush=True)
    for flag in ("0", "1"):
        os.environ["QCK_SIM_TMA"] = flag
        print("running TMA =", flag, flush=True)
        r = eb.run(h); torch.cuda.synchronize()
        print("  max err vs on-chip", np.abs(r.cpu().numpy() - ra).max(), flush=True)
