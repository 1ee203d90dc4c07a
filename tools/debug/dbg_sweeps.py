"""Scratch: state after each sweep, plain kernel vs numpy interpreter (tests/plan_interpreter.py)."""
import os, sys, copy, ctypes as C
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from importlib import import_module
import plan_interpreter as pi
PKG = "hardwareawareoptimalquantumcircuitcuttingandknitting_b200"
cutting = import_module(f"{PKG}.cutting"); vcm = import_module(f"{PKG}.virtual_circuit")
compiler = import_module(f"{PKG}.compiler"); _lib = import_module(f"{PKG}._lib")
cfg, onchip, tile = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda", 0)
circ, cut = cutting.make_baseline(cfg)
virt = vcm.VirtualCircuit(cut)
h = _lib.get_handle(0)
stream = torch.cuda.current_stream(dev).cuda_stream
for f in virt.active_fragments():
    b = compiler.FragmentProgram(virt.fragment_circuits[f], f, virt.num_clbits, onchip_max=onchip, stream_tile=tile)
    eb = compiler.FragmentExecutor(b, dev); eb.upload()
    for pi_idx, plan in enumerate(eb.plans):
        st = eb.plan_struct(pi_idx)
        label = int(plan.labels[0])
        n_sw = len(plan.sweeps)
        for k in range(1, n_sw + 1):
            trunc = copy.copy(plan); trunc.sweeps = plan.sweeps[:k]
            want = pi.run_plan(b, trunc, label, return_state=True)
            st.n_sweeps = k
            for flag in ("0", "1"):
                os.environ["QCK_SIM_TMA"] = flag
                buf = torch.zeros(2 << plan.n_state, dtype=torch.float64, device=dev)
                h.check(h.lib.qck_sim_statevector(h.ptr, C.byref(st), label, buf.data_ptr(), buf.numel() * 8, stream))
                got = buf.cpu().numpy().view(np.complex128)
                err = np.abs(got - want)
                print(f"frag {len(f)}q plan {pi_idx} n_state {plan.n_state} sweeps<= {k}/{n_sw} tile {plan.sweeps[k-1][0]} TMA={flag}: max err {err.max():.3e}"
                      + (f" first bad amp {int(np.argmax(err > 1e-12))}" if err.max() > 1e-12 else ""), flush=True)
        st.n_sweeps = n_sw
        break
