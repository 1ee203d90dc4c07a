"""Uncut statevector sharded over the GPUs of one box (torchrun, one rank per GPU).

usage: python -m torch.distributed.run --nproc-per-node G tools/sharded_statevector.py [name=syc] [n=33] [depth=1] [reps=3]
Prints one JSON line from rank 0: device time (CUDA events, max over ranks), norm, exact bytes moved and the
share that crossed NVLink; for n <= 29 every rank also checks its shard against a single-GPU run.
"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.distributed as dist
from importlib import import_module

PKG = "hardwareawareoptimalquantumcircuitcuttingandknitting_b200"
gen = import_module(PKG + ".generators")
sharded = import_module(PKG + ".sharded")
_lib = import_module(PKG + "._lib")

name = sys.argv[1] if len(sys.argv) > 1 else "syc"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 33
depth = int(sys.argv[3]) if len(sys.argv) > 3 else 1
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

circ = gen.gen_circ(name, n, depth, seed=0).decompose_two_qubit()
sv = sharded.ShardedStatevector(circ, dev, rank=rank, world=world)
sv.run()                                                     # warm-up
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    sv.run()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
norm = sv.norm()
err = None
if n <= 29:                                                  # every rank checks its shard against a single-GPU run
    ex = sv.ex
    st = ex.plan_struct(0)
    h = _lib.get_handle(local)
    ref = torch.empty(2 << n, dtype=torch.float64, device=dev)
    h.check(h.lib.qck_sim_statevector(h.ptr, C.byref(st), 0, ref.data_ptr(), ref.numel() * 8,
                                      torch.cuda.current_stream(dev).cuda_stream))
    torch.cuda.synchronize()
    m = 2 << sv.n_local
    e = (sv.local_shard() - ref[rank * m:(rank + 1) * m]).abs().max().reshape(1)
    if world > 1:
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
    err = float(e.item())
tr = sv.traffic()
if rank == 0:
    sweeps = sv.plan.sweeps
    print(json.dumps({"circuit": f"{name}-{n} d{depth}", "n_gpus": world, "local_qubits": sv.n_local, "sweeps": len(sweeps),
                      "sweeps_touching_peers": sum(1 for s in range(len(sweeps)) if sv._touches_peers(s)),
                      "ms": float(ms.item()), "norm": norm, "max_abs_err_vs_single_gpu": err,
                      "bytes_moved": tr["bytes"], "bytes_over_nvlink": tr["peer_bytes"],
                      "aggregate_gbs": tr["bytes"] / float(ms.item()) / 1e6,
                      "nvlink_gbs_per_gpu": tr["peer_bytes"] / world / float(ms.item()) / 1e6}))
sv.close()
if world > 1:
    dist.destroy_process_group()
