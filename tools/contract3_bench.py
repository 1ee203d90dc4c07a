"""Label contraction for a three-fragment cut (chain A - B - C, 16 output bits split 6 | 5 | 5, three cx cuts per
boundary: 6^6 = 46 656 labels): grouped tensor-core path against the per-output generic kernel (CUDA events)."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from importlib import import_module

PKG = "hardwareawareoptimalquantumcircuitcuttingandknitting_b200"
_lib = import_module(f"{PKG}._lib")
dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
radices = [6] * 6
masks = [0x003F, 0x07C0, 0xF800]
touches = [[True] * 3 + [False] * 3, [True] * 6, [False] * 3 + [True] * 3]
n_out, F, K = 16, 3, 6
tables = [rng.normal(size=(int(np.prod([r for r, t in zip(radices, tt) if t])), 1 << bin(m).count("1")))
          for m, tt in zip(masks, touches)]
h = _lib.get_handle(0)
d_t = [torch.from_numpy(t).to(dev) for t in tables]
ptrs = (C.c_void_p * F)(*[t.data_ptr() for t in d_t])
cm = (C.c_uint64 * F)(*masks)
rs = (C.c_int64 * F)(*[t.shape[1] for t in tables])
rad = (C.c_int32 * K)(*radices)
coef = (C.c_double * (K * _lib.MAX_VARIANTS))(*rng.normal(size=K * _lib.MAX_VARIANTS))
st = (C.c_int32 * (F * _lib.MAX_DIGITS))()
for f in range(F):
    acc = 1
    for k in reversed(range(K)):
        if touches[f][k]:
            st[f * _lib.MAX_DIGITS + k] = acc
            acc *= radices[k]
out = torch.empty(1 << n_out, dtype=torch.float64, device=dev)
stream = torch.cuda.current_stream(dev).cuda_stream
total = int(np.prod(radices))
res = {}
for mode in ("1", "0"):
    os.environ["QCK_CONTRACT_GROUPS"] = mode
    for _ in range(3):
        h.check(h.lib.qck_knit_contract(h.ptr, F, ptrs, cm, rs, n_out, K, rad, coef, st, 0, total, out.data_ptr(), 0, stream))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        h.check(h.lib.qck_knit_contract(h.ptr, F, ptrs, cm, rs, n_out, K, rad, coef, st, 0, total, out.data_ptr(), 0, stream))
    e1.record()
    torch.cuda.synchronize()
    res["grouped" if mode == "1" else "generic"] = {"ms": e0.elapsed_time(e1) / 10, "vec": out.clone()}
err = (res["grouped"]["vec"] - res["generic"]["vec"]).abs().max().item() / res["generic"]["vec"].abs().max().item()
flop = 2.0 * total * (1 << n_out)
print(json.dumps({"labels": total, "n_out": n_out, "fragments": "6|5|5 bits", "grouped_ms": res["grouped"]["ms"],
                  "generic_ms": res["generic"]["ms"], "rel_diff": err,
                  "grouped_tflops_2L2^n": flop / res["grouped"]["ms"] / 1e9}))
