"""Uncut statevector sharded over the GPUs of one box.

The reference scores a cut run against the ideal run of the UNCUT circuit
(``src/HwAwareCutter/Utilities.py:39-69``, Aer on the CPU).  One B200 holds 32 qubits (64 GiB of
complex128); beyond that the state is split by its top ``log2(world)`` index bits, one shard per rank.
There is no exchange step: every rank runs the same sweeps on the tiles it owns, and a sweep whose tile
contains rank bits loads / stores the peer halves of its tiles straight from / to the peers' buffers with
TMA over NVLink (``qck_sim_sweeps_sharded``; buffers mapped through CUDA IPC).  Between sweeps the ranks
are ordered with a stream-ordered one-element NCCL all-reduce, only where a sweep touches peer memory.

``emulate=True`` keeps every shard on ONE device and runs the ranks one after the other on one stream
(same kernels, same tensor maps, no IPC): used by the single-GPU tests.
"""
from __future__ import annotations

import ctypes as C

from . import _lib
from .compiler import FragmentExecutor, FragmentProgram
from .virtual_circuit import VirtualCircuit


class _RawBuffer:
    """Device memory from ``qck_mem_alloc`` (plain cudaMalloc: exportable through CUDA IPC) with the CUDA
    array interface, so that ``torch.as_tensor`` can view it."""

    def __init__(self, handle: "_lib.Handle", nbytes: int) -> None:
        self.handle = handle
        self.nbytes = int(nbytes)
        ptr = C.c_void_p()
        handle.check(handle.lib.qck_mem_alloc(handle.ptr, self.nbytes, C.byref(ptr)))
        self.ptr = ptr.value

    @property
    def __cuda_array_interface__(self) -> dict:
        return {"shape": (self.nbytes // 8,), "typestr": "<f8", "data": (self.ptr, False), "version": 2}

    def free(self) -> None:
        if self.ptr:
            self.handle.lib.qck_mem_free(self.handle.ptr, C.c_void_p(self.ptr))
            self.ptr = None


class ShardedStatevector:
    """Final statevector of an uncut circuit, ``2^(n - log2(world))`` amplitudes per rank."""

    def __init__(self, circ, device, rank: int = 0, world: int = 1, emulate: bool = False) -> None:
        import torch
        self.torch = torch
        self.device = torch.device(device)
        self.rank, self.world, self.emulate = int(rank), int(world), bool(emulate)
        if self.world < 1 or self.world & (self.world - 1) or self.world > 8:
            raise ValueError("world must be a power of two <= 8")
        self.virt = VirtualCircuit(circ)
        (frag,) = self.virt.active_fragments()
        self.g = self.world.bit_length() - 1
        n = len(frag)
        self.n = n
        self.n_local = n - self.g
        if self.n_local < 14:
            raise ValueError(f"{self.n} qubits over {self.world} ranks leaves fewer than 14 local qubits")
        # Two schedules: the plain one, and one whose FIRST sweep takes the rank bits into its tile - they then
        # go live while the state is a single tile and the large expansion sweeps of a shallow circuit write
        # local memory only (syc-33 d1 on 2 GPUs: 68.7 GB -> 0 GB over NVLink); deeper circuits can be better
        # off with the plain one.  The one that moves fewer bytes between shards wins (host arithmetic).
        early = ((1 << n) - 1) & ~((1 << self.n_local) - 1)
        best = None
        for bits in ((early, 0) if self.world > 1 else (0,)):
            prog = FragmentProgram(self.virt.fragment_circuits[frag], frag, self.virt.num_clbits, early_bits=bits)
            if prog.radix:
                raise ValueError("sharded runs simulate one uncut circuit (no virtual gates)")
            ex = FragmentExecutor(prog, self.device, True)
            if len(ex.plans) != 1 or ex.plans[0].n_state != n:
                raise ValueError("sharded runs need a measurement-terminal circuit (no ancilla bits)")
            self.ex, self.plan = ex, ex.plans[0]
            tr = self.traffic()
            key = (tr["peer_bytes"], tr["bytes"])
            if best is None or key < best[0]:
                best = (key, ex)
        self.ex, self.plan = best[1], best[1].plans[0]
        self.handle = _lib.get_handle(self.device.index or 0)
        self.shard_bytes = 16 << self.n_local
        self._own: list = []          # buffers this process allocated (raw or torch)
        self._opened: list = []       # peer pointers opened through IPC
        self.ptrs: list = []          # device pointers of all shards, by rank
        self._token = None
        self._allocate()

    # ------------------------------------------------------------------ memory
    def _allocate(self) -> None:
        torch, h = self.torch, self.handle
        if self.emulate or self.world == 1:
            for _ in range(self.world):
                buf = torch.empty(self.shard_bytes // 8, dtype=torch.float64, device=self.device)
                self._own.append(buf)
                self.ptrs.append(buf.data_ptr())
            return
        import torch.distributed as dist
        raw = _RawBuffer(h, self.shard_bytes)
        self._own.append(raw)
        ipc = C.create_string_buffer(64)
        h.check(h.lib.qck_ipc_export(h.ptr, C.c_void_p(raw.ptr), ipc))
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(ipc.raw))
        for r, hb in enumerate(handles):
            if r == self.rank:
                self.ptrs.append(raw.ptr)
                continue
            p = C.c_void_p()
            h.check(h.lib.qck_ipc_open(h.ptr, hb, C.byref(p)))
            self._opened.append(p.value)
            self.ptrs.append(p.value)
        self._token = torch.zeros(1, dtype=torch.float32, device=self.device)

    def close(self) -> None:
        h = self.handle
        if not self.emulate and self.world > 1:
            import torch.distributed as dist
            self.torch.cuda.synchronize(self.device)
            dist.barrier()                       # nobody unmaps / frees while a peer may still use the memory
        for p in self._opened:
            h.lib.qck_ipc_close(h.ptr, C.c_void_p(p))
        self._opened = []
        for b in self._own:
            if isinstance(b, _RawBuffer):
                b.free()
        self._own = []

    # ------------------------------------------------------------------ run
    def _touches_peers(self, s: int) -> bool:
        return any(p >= self.n_local for p in self.plan.sweeps[s][0])

    def _order_ranks(self) -> None:
        """All ranks have finished what they enqueued so far before any continues (stream ordered)."""
        if self.emulate or self.world == 1:
            return
        import torch.distributed as dist
        dist.all_reduce(self._token)

    def run(self) -> None:
        """Enqueue every sweep (asynchronous on the current stream)."""
        h = self.handle
        st = self.ex.plan_struct(0)
        stream = self.torch.cuda.current_stream(self.device).cuda_stream
        ptrs = (C.c_void_p * self.world)(*self.ptrs)
        n_sw = len(self.plan.sweeps)
        ranks = range(self.world) if self.emulate else (self.rank,)
        s = 0
        while s < n_sw:
            e = s + 1
            if not self._touches_peers(s):       # a run of purely local sweeps goes out in one call
                while e < n_sw and not self._touches_peers(e):
                    e += 1
            else:
                self._order_ranks()              # the peers' previous sweeps wrote what this one reads
            for r in ranks:
                h.check(h.lib.qck_sim_sweeps_sharded(h.ptr, C.byref(st), s, e, r, self.world, ptrs, self.shard_bytes,
                                                     stream))
            if self._touches_peers(s):
                self._order_ranks()              # this sweep wrote into the peers' shards
            s = e

    # ------------------------------------------------------------------ results
    def local_shard(self, rank: int | None = None):
        """float64 view [2 * 2^n_local] (re, im interleaved) of a shard held by this process."""
        torch = self.torch
        if self.emulate or self.world == 1:
            return self._own[self.rank if rank is None else rank]
        return torch.as_tensor(self._own[0], device=self.device)

    def norm(self) -> float:
        """Sum of |amp|^2 over the whole state (all-reduced over the ranks)."""
        torch = self.torch
        if self.emulate or self.world == 1:
            return float(sum((b * b).sum() for b in self._own))
        import torch.distributed as dist
        v = self.local_shard()
        t = (v * v).sum().reshape(1)
        dist.all_reduce(t)
        return float(t.item())

    def traffic(self) -> dict:
        """Exact bytes this configuration moves (host arithmetic over the sweep descriptions): total and the
        part that crosses between shards."""
        lib = _lib.load()
        st = _lib.QckSimPlan.from_buffer_copy(self.ex._structs[0][0])      # host fields only: no device pointers needed
        geom, perm = (C.c_int32 * 8)(), (C.c_int32 * 16)()
        ld_off, st_off = (C.c_uint64 * 128)(), (C.c_uint64 * 128)()
        ld_slot, st_slot = (C.c_uint32 * 128)(), (C.c_uint32 * 128)()
        em, nw, fx = C.c_uint64(), C.c_uint64(), C.c_uint64()
        live, total, peer = 0, 0, 0
        n_sw = len(self.plan.sweeps)
        for i, (positions, _b, _e) in enumerate(self.plan.sweeps):
            T = len(positions)
            for r in range(self.world):
                ok = lib.qck_debug_tma_describe(C.byref(st), i, live, int(i == n_sw - 1), 1,
                                                self.n_local if self.world > 1 else 0, r, geom, perm, ld_off, ld_slot,
                                                st_off, st_slot, C.byref(em), C.byref(nw), C.byref(fx))
                if not ok:
                    raise NotImplementedError(f"sweep {i} is not eligible for the TMA kernel")
                n_load, n_store, zf_shift = geom[3], geom[4], geom[5]
                fixed = fx.value
                live_tiles = (1 << bin(em.value & live).count("1")) if (fixed & ~live) == 0 else 0
                if nw.value == 0:
                    continue
                ld_b = live_tiles * n_load * (16 << zf_shift)
                st_b = nw.value * (16 << T)
                total += ld_b + st_b
                if self.world > 1:
                    in_tile = [p for p in positions if p >= self.n_local]
                    if in_tile:                   # boxes whose rank bits differ from the working rank's
                        def frac(offs, cnt):
                            other = sum(1 for j in range(cnt) if ((fixed | offs[j]) >> self.n_local) != r)
                            return other / cnt if cnt else 0.0
                        peer += int(ld_b * frac(ld_off, n_load) + st_b * frac(st_off, n_store))
            for p in positions:
                live |= 1 << p
        return {"bytes": total, "peer_bytes": peer}
