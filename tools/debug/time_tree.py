"""Scratch: where the time of the tree-walk simulation goes (QCK_TREE_DEBUG_SKIP: 1 = no combine, 2 = no last
level, 4 = only the last level)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
from importlib import import_module
PKG = "hardwareawareoptimalquantumcircuitcuttingandknitting_b200"
cutting = import_module(PKG + ".cutting"); vcm = import_module(PKG + ".virtual_circuit"); lib = import_module(PKG + "._lib")
wl = sys.argv[1] if len(sys.argv) > 1 else "hwe16d5"
dev = torch.device("cuda", 0)
circ, cut = cutting.make_baseline(wl, 0)
virt = vcm.VirtualCircuit(cut)
h = lib.get_handle(0)
frags = virt.active_fragments()
exs = [virt.executor(f, dev, True) for f in frags]
outs = [ex.run(h) for ex in exs]
torch.cuda.synchronize()

def timed(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

print(wl, "fragment 0 alone  : %.1f us" % timed(lambda: exs[0].run(h, out=outs[0])))
print(wl, "fragment 1 alone  : %.1f us" % timed(lambda: exs[1].run(h, out=outs[1])))
print(wl, "both, one after the other: %.1f us" % timed(lambda: [ex.run(h, out=o) for ex, o in zip(exs, outs)]))
tabs = dict(zip(frags, outs))
print(wl, "both, in a region : %.1f us" % timed(lambda: virt.simulate_fragments(dev, out=tabs)))
