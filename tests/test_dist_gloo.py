"""N > 1 host logic on CPU: world_size-2 gloo processes (the compute is the oracle's, the
partitioning / reduction plumbing is the product's dist.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "hardwareawareoptimalquantumcircuitcuttingandknitting_b200"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    from importlib import import_module
    qdist = import_module(f"{PKG}.dist")
    from oracle import dense as od
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    assert qdist.init_from_env("gloo") == (rank, rank, world)
    rng = np.random.default_rng(7)
    tA, tB = rng.random(1 << 6), rng.random(1 << 3)
    mA, mB = 0b101101011, 0b010010100
    n_out = 9
    # (a) no virtual gates: output sharded by the top bit, scalar all-reduce only
    y0, y1 = qdist.shard_pow2(n_out, rank, world)
    part = od.knit_outer([tA, tB], [mA, mB], y0, y1)
    stats = torch.tensor([part.sum(), part.min(), 0.0, float(np.count_nonzero(part))], dtype=torch.float64)
    qdist.allreduce_stats(stats)
    full = od.knit_outer([tA, tB], [mA, mB], 0, 1 << n_out)
    assert abs(stats[0].item() - full.sum()) < 1e-12 and stats[1].item() == full.min()
    assert stats[3].item() == np.count_nonzero(full)
    # (b) virtual gates: labels sharded (aligned to the innermost radix), dense all-reduce
    L, radix_last = 36, 6
    lo, hi = qdist.shard_range(L, rank, world, align=radix_last)
    assert lo % radix_last == 0 and (hi % radix_last == 0 or hi == L)
    terms = rng.random((L, 1 << 6))
    partial = torch.from_numpy(terms[lo:hi].sum(axis=0))
    qdist.allreduce_sum_(partial)
    assert np.abs(partial.numpy() - terms.sum(axis=0)).max() < 1e-12
    ret[rank] = (y0, y1, lo, hi)
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo():
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret[0][:2] == (0, 256) and ret[1][:2] == (256, 512)
    assert ret[0][2:] == (0, 18) and ret[1][2:] == (18, 36)


def test_shard_helpers():
    sys.path.insert(0, ROOT)
    from importlib import import_module
    qdist = import_module(f"{PKG}.dist")
    for total, world, align in [(7776, 8, 6), (8, 4, 8), (1296, 3, 6), (5, 8, 1)]:
        spans = [qdist.shard_range(total, r, world, align) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == total
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    assert [qdist.shard_pow2(32, r, 8)[0] >> 29 for r in range(8)] == list(range(8))
    with pytest.raises(ValueError):
        qdist.shard_pow2(4, 0, 3)
