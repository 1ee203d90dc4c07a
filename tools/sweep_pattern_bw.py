"""Scratch: bandwidth of an EMPTY sweep (load tile, store tile) as a function of the tile shape."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from importlib import import_module
PKG = "hardwareawareoptimalquantumcircuitcuttingandknitting_b200"
_lib = import_module(PKG + "._lib")
h = _lib.get_handle(0)
dev = torch.device("cuda", 0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 30
state = torch.zeros(2 << N, dtype=torch.float64, device=dev)
ops = torch.zeros(8, dtype=torch.int32, device=dev); mats = torch.zeros(8, dtype=torch.float64, device=dev)
label = torch.zeros(1, dtype=torch.int32, device=dev)

def run(T, low, hi_start):
    pos = list(range(low)) + list(range(hi_start, hi_start + T - low))
    sw = (_lib.QckSweep * 2)()
    for k in range(2):
        sw[k].n_tile = T; sw[k].op_begin = 0; sw[k].op_end = 0
        for j, x in enumerate(pos): sw[k].pos[j] = x
    plan = _lib.QckSimPlan()
    plan.n_state_qubits = N; plan.n_sweeps = 2; plan.sweeps = sw
    plan.d_ops = ops.data_ptr(); plan.d_mats = mats.data_ptr(); plan.n_digits = 0
    plan.n_out_bits = 0; plan.sum_mask = (1 << N) - 1; plan.sign_mask = 0
    f = lambda: h.check(h.lib.qck_sim_statevector(h.ptr, C.byref(plan), 0, state.data_ptr(), state.numel() * 8, 0))
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); [f() for _ in range(3)]; e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    byts = (16 << N) * 3          # sweep 0: init (write only), sweep 1: read + write
    return ms, byts / ms / 1e6

for pipe in ("0", "1"):
    os.environ["QCK_SIM_PIPE"] = pipe
    for T in (11, 12):
        for low in (5, 6, 7, 8):
            for hi_start in (13, N - (T - low)):
                ms, gbs = run(T, low, hi_start)
                print(f"pipe={pipe} T={T} low_run={low} ({16<<low} B runs) hi bits from {hi_start}: {ms:.2f} ms  {gbs:.0f} GB/s")
