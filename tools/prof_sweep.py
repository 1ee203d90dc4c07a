"""Uncut statevector run for profiling the streaming kernels.

usage: prof_sweep.py [name=syc] [n=28] [depth=1] [reps=3]      (QCK_SIM_TMA=0 selects the plain kernel)
Prints the time of the sweeps alone (qck_sim_statevector) and of the whole fragment run (sweeps + fold),
and the bandwidth the sweeps would need if every sweep read and wrote the whole state (the plain
kernel's traffic; the TMA path with live-qubit tracking moves less on shallow circuits).
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from importlib import import_module

PKG = "hardwareawareoptimalquantumcircuitcuttingandknitting_b200"
gen = import_module(PKG + ".generators")
vcm = import_module(PKG + ".virtual_circuit")
_lib = import_module(PKG + "._lib")
if os.environ.get("QCK_LIB_VARIANT"):            # A/B experiments: load libqck_<variant>.so instead
    _lib.LIB_PATH = _lib.LIB_PATH.replace("libqck.so", "libqck_" + os.environ["QCK_LIB_VARIANT"] + ".so")

name = sys.argv[1] if len(sys.argv) > 1 else "syc"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 28
depth = int(sys.argv[3]) if len(sys.argv) > 3 else 1
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
circ = gen.gen_circ(name, n, depth, seed=0).decompose_two_qubit()
virt = vcm.VirtualCircuit(circ)
(frag,) = virt.active_fragments()
dev = torch.device("cuda", 0)
compiler = import_module(PKG + ".compiler")
if os.environ.get("QCK_PROF_CLUSTER", "1") == "0":     # experiment: every op as its own shared-memory pass
    prog = compiler.FragmentProgram(virt.fragment_circuits[frag], frag, circ.num_clbits, cluster=False)
    ex = compiler.FragmentExecutor(prog, dev, True)
else:
    ex = virt.executor(frag, dev, True)
ex.upload()
h = _lib.get_handle(0)
pl = ex.plans[0]
prog_radix = list(ex.program.radix)
st = ex.plan_struct(0)
stream = torch.cuda.current_stream(dev).cuda_stream
state = torch.empty(2 << n, dtype=torch.float64, device=dev)


def sweeps_only():
    h.check(h.lib.qck_sim_statevector(h.ptr, C.byref(st), 0, state.data_ptr(), state.numel() * 8, stream))


sweeps_only()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    sweeps_only()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
n_sw = len(pl.sweeps)
full_bytes = (2 * n_sw - 1) * (16 << n)
ld, sd, used = C.c_uint64(), C.c_uint64(), C.c_int()
h.check(h.lib.qck_sim_plan_traffic(C.byref(st), 1, 0, C.byref(ld), C.byref(sd), C.byref(used)))
moved = ld.value + sd.value
norm = float((state * state).sum())
print(f"{name}-{n} d{depth} uncut, QCK_SIM_TMA={os.environ.get('QCK_SIM_TMA', '1')}: sweeps {n_sw}, records {len(pl.ops)}, "
      f"sweeps-only {ms:.3f} ms = {full_bytes / ms / 1e6:.0f} GB/s of full-sweep traffic ({full_bytes / 1e9:.1f} GB), "
      f"norm {norm:.12f}")
print(f"  HBM bytes moved by the sweeps ({'TMA path, live-qubit tracking' if used.value else 'plain path'}): "
      f"{ld.value / 1e9:.2f} GB loaded + {sd.value / 1e9:.2f} GB stored = {moved / ms / 1e6:.0f} GB/s")
del state
t = ex.run(h)
torch.cuda.synchronize()
e0.record()
for _ in range(reps):
    ex.run(h, out=t)
e1.record()
torch.cuda.synchronize()
ms_all = e0.elapsed_time(e1) / reps
fused = (used.value and os.environ.get("QCK_FOLD_FUSION", "1") != "0" and not prog_radix
         and list(pl.out_pos) == list(range(pl.n_state)) and pl.sum_mask == 0)
h.check(h.lib.qck_sim_plan_traffic(C.byref(st), 1, int(bool(fused)), C.byref(ld), C.byref(sd), C.byref(used)))
tot = ld.value + sd.value + (0 if fused else (16 << n) + (8 << len(pl.out_pos)))
print(f"  whole fragment run (sweeps + {'fused fold' if fused else 'fold pass'}): {ms_all:.3f} ms, {tot / 1e9:.2f} GB = "
      f"{tot / ms_all / 1e6:.0f} GB/s, sum {t.sum().item():.12f}")
for pos, b, e in pl.sweeps:
    print("  sweep tile", pos, "records", e - b)
