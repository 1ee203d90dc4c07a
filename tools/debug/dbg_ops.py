"""Scratch: first sweep only, ops[0:m] for growing m, plain kernel vs interpreter."""
import os, sys, copy, ctypes as C
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from importlib import import_module
import plan_interpreter as pi
PKG = "hardwareawareoptimalquantumcircuitcuttingandknitting_b200"
cutting = import_module(f"{PKG}.cutting"); vcm = import_module(f"{PKG}.virtual_circuit")
compiler = import_module(f"{PKG}.compiler"); _lib = import_module(f"{PKG}._lib")
if os.environ.get("QCK_LIB_VARIANT"):
    _lib.LIB_PATH = _lib.LIB_PATH.replace("libqck.so", "libqck_" + os.environ["QCK_LIB_VARIANT"] + ".so")
cfg, onchip, tile = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda", 0)
circ, cut = cutting.make_baseline(cfg)
virt = vcm.VirtualCircuit(cut)
h = _lib.get_handle(0)
stream = torch.cuda.current_stream(dev).cuda_stream
f = virt.active_fragments()[0]
b = compiler.FragmentProgram(virt.fragment_circuits[f], f, virt.num_clbits, onchip_max=onchip, stream_tile=tile)
eb = compiler.FragmentExecutor(b, dev); eb.upload()
plan = eb.plans[0]
st = eb.plan_struct(0)
label = int(plan.labels[0])
pos, b0, e0 = plan.sweeps[0]
st.n_sweeps = 1
os.environ["QCK_SIM_TMA"] = "0"
for m in range(1, e0 - b0 + 1):
    trunc = copy.copy(plan); trunc.sweeps = [(pos, b0, b0 + m)]
    want = pi.run_plan(b, trunc, label, return_state=True)
    st.sweeps[0].op_end = st.sweeps[0].op_begin + m
    buf = torch.zeros(2 << plan.n_state, dtype=torch.float64, device=dev)
    h.check(h.lib.qck_sim_statevector(h.ptr, C.byref(st), label, buf.data_ptr(), buf.numel() * 8, stream))
    got = buf.cpu().numpy().view(np.complex128)
    err = np.abs(got - want)
    print(f"ops[0:{m}] last {plan.ops[b0 + m - 1].tolist()} max err {err.max():.3e}", (f"first bad {int(np.argmax(err > 1e-12))} got {got[int(np.argmax(err > 1e-12))]} want {want[int(np.argmax(err > 1e-12))]}" if err.max() > 1e-12 else ""), flush=True)
