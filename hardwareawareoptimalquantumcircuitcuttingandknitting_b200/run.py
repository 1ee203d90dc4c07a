"""run_virtual_circuit: instantiate -> execute -> knit -> nearest distribution.

Mirror of ``third_party/qvm/qvm/run.py:17-71`` with the reference signature
``run_virtual_circuit(virt, shots=20000) -> (dict[int, float], RunTimeInfo)``.

When every fragment is bound to a ``B200Backend`` (the default) the whole path
runs on the GPU from the compiled templates: exact outcome distributions instead
of ``shots`` samples (``shots`` is accepted and ignored), signed-folded fragment
tables, closed-form knit, ``nearest_probability_distribution`` on the dense
vector.  ``run_virtual_circuit_dense`` is the same path returning the dense
device tensor - the only usable form at 32 output bits, where a Python dict of
2^32 entries cannot exist.

If a fragment is bound to a foreign duck-typed backend
(``.run(circuits, shots=) -> job``, ``job.result().get_counts()``; ``run.py:42,
48-56``) the reference's flow is followed literally: host-side instance
circuits, the backend's counts, ``QuasiDistr.from_counts`` and the
reference-order ``virt.knit`` - still on device-resident distributions.

The four log lines of the reference (``run.py:28-32,45,62,69``) are kept, through
the standard ``logging`` module.
"""
from __future__ import annotations

import ctypes as C
import logging
from dataclasses import dataclass
from time import perf_counter

import numpy as np

from . import _lib
from .backend import B200Backend
from .quasi_distr import QuasiDistr, default_device
from .virtual_circuit import VirtualCircuit, generate_instantiations

logger = logging.getLogger(__name__)

__all__ = ["RunTimeInfo", "DenseResult", "run_virtual_circuit", "run_virtual_circuit_dense"]


@dataclass
class RunTimeInfo:
    run_time: float
    knit_time: float


@dataclass
class DenseResult:
    """Dense full-circuit distribution on the device.

    ``values[i]``: probability of the bitstring whose written clbits, taken in ascending
    order, spell ``i`` (``key_mask`` = the written clbits; for ``measure_all`` circuits
    ``i`` is the key itself).  ``total`` / ``minimum`` are the sum and the smallest entry of
    the knitted quasi-distribution before ``nearest_probability_distribution``."""
    values: object
    key_mask: int
    total: float
    minimum: float
    y_begin: int = 0

    def to_dict(self) -> dict[int, float]:
        host = self.values.cpu().numpy()
        nz = np.nonzero(host)[0]
        keys = _pdep_array(nz.astype(np.uint64) + np.uint64(self.y_begin), self.key_mask)
        return {int(k): float(host[i]) for k, i in zip(keys.tolist(), nz.tolist())}


def _pdep_array(x: np.ndarray, mask: int) -> np.ndarray:
    if mask == (1 << mask.bit_length()) - 1:
        return x
    out = np.zeros_like(x)
    j = 0
    for b in range(mask.bit_length()):
        if (mask >> b) & 1:
            out |= ((x >> np.uint64(j)) & np.uint64(1)) << np.uint64(b)
            j += 1
    return out


def _all_b200(virt: VirtualCircuit) -> bool:
    return all(isinstance(virt.get_backend(f), B200Backend) for f in virt.fragment_circuits)


def run_virtual_circuit_dense(virt: VirtualCircuit, shots: int = 20000, device=None, nearest: bool = True,
                              rank: int = 0, world_size: int = 1, group=None,
                              out=None, accuracy: float | None = None) -> tuple[DenseResult, RunTimeInfo]:
    """Whole path on the GPU.  With ``world_size > 1`` (one process per GPU) the work is
    partitioned as described in ``dist.py``: without virtual gates this rank produces the
    output slice whose top bits equal ``rank`` (``DenseResult.y_begin`` tells where it
    starts); with virtual gates the labels are sharded and the dense result is all-reduced,
    so every rank returns the full vector."""
    import torch
    from . import dist as qdist
    if not _all_b200(virt):
        raise ValueError("run_virtual_circuit_dense needs every fragment on a B200Backend")
    device = default_device() if device is None else torch.device(device)
    handle = _lib.get_handle(device.index or 0)
    frags = virt.fragment_circuits
    logger.info(f"Running virtualizer with {len(frags)} "
                + f"{tuple(circ.num_qubits for circ in frags.values())} "
                + f"fragments and {len(virt._vgate_instrs)} vgates...")
    K = len(virt._vgate_instrs)
    from . import quasi_distr as _qd
    accuracy = _qd.ACCURACY if accuracy is None else float(accuracy)
    if accuracy > 0.0:
        return _run_faithful(virt, device, handle, nearest, accuracy, world_size, out)
    label_range = None
    if K > 0 and world_size > 1:
        label_range = qdist.shard_range(virt.num_global_labels(), rank, world_size,
                                        align=virt.global_radices()[-1])
    now = perf_counter()
    logger.info(f"Running {sum(virt.program(f).num_labels for f in virt.active_fragments())} instances...")
    tables = virt.simulate_fragments(device, label_range=label_range)
    run_time = perf_counter() - now          # enqueue time; the device work is timed with the knit

    logger.info("Knitting...")
    now = perf_counter()
    stats = torch.zeros(4, dtype=torch.float64, device=device)
    _, union = virt.output_masks()
    y_begin = 0
    if K == 0:
        n_out = bin(union).count("1")
        y_range = qdist.shard_pow2(n_out, rank, world_size) if world_size > 1 else None
        y_begin = y_range[0] if y_range else 0
        values = virt.knit_tables(tables, device, stats=stats, y_range=y_range, out=out)
        qdist.allreduce_stats(stats, group)
    else:
        values = virt.knit_tables(tables, device, label_range=label_range, out=out,
                                  stats=None if world_size > 1 else stats)
        if world_size > 1:
            qdist.allreduce_sum_(values, group)
            stream = torch.cuda.current_stream(device).cuda_stream
            handle.check(handle.lib.qck_stats_dense(handle.ptr, values.data_ptr(), values.numel(), 0.0,
                                                    stats.data_ptr(), stream))
    host_stats = stats.cpu().numpy()          # the step's device -> host read (synchronises)
    total, minimum = float(host_stats[0]), float(host_stats[1])
    if nearest and minimum < 0.0:
        if K == 0 and world_size > 1:
            raise NotImplementedError("nearest_probability_distribution on an output-sharded result")
        stream = torch.cuda.current_stream(device).cuda_stream
        handle.check(handle.lib.qck_npd(handle.ptr, values.data_ptr(), values.numel(), 0.0, None, None, stream))
    torch.cuda.synchronize(device)
    knit_time = perf_counter() - now
    logger.info(f"Knitted in {knit_time:.2f}s.")
    return DenseResult(values, union, total, minimum, y_begin), RunTimeInfo(run_time, knit_time)


def _run_faithful(virt, device, handle, nearest, accuracy, world_size, out):
    """ACCURACY > 0: exact instance distributions, then the reference's pruned algebra in the
    reference's order, fused per output entry (``qck_knit_faithful``).  Single GPU."""
    import torch
    if world_size > 1:
        raise NotImplementedError("the reference-faithful knit is not sharded")
    now = perf_counter()
    tables = virt.simulate_fragments(device, fold=False)
    run_time = perf_counter() - now
    logger.info("Knitting...")
    now = perf_counter()
    stats = torch.zeros(4, dtype=torch.float64, device=device)
    values = virt.knit_tables_faithful(tables, accuracy, device, out=out, stats=stats)
    host_stats = stats.cpu().numpy()
    total, minimum = float(host_stats[0]), float(host_stats[1])
    if nearest and host_stats[3] > 0 and minimum < 0.0:
        stream = torch.cuda.current_stream(device).cuda_stream
        handle.check(handle.lib.qck_npd(handle.ptr, values.data_ptr(), values.numel(), accuracy, None, None, stream))
    torch.cuda.synchronize(device)
    knit_time = perf_counter() - now
    logger.info(f"Knitted in {knit_time:.2f}s.")
    _, union = virt.output_masks()
    return DenseResult(values, union, total, minimum, 0), RunTimeInfo(run_time, knit_time)


def run_virtual_circuit(virt: VirtualCircuit, shots: int = 20000) -> tuple[dict[int, float], RunTimeInfo]:
    if _all_b200(virt):
        dense, info = run_virtual_circuit_dense(virt, shots)
        if dense.values.numel() > (1 << 26):
            raise MemoryError("result too wide for a Python dict; use run_virtual_circuit_dense")
        return dense.to_dict(), info

    # ---- foreign backends: the reference's flow, on device-resident distributions
    jobs = {}
    frags = virt.fragment_circuits
    logger.info(f"Running virtualizer with {len(frags)} "
                + f"{tuple(circ.num_qubits for circ in frags.values())} "
                + f"fragments and {len(virt._vgate_instrs)} vgates...")
    num_instances = 0
    now = perf_counter()
    for frag, frag_circuit in frags.items():
        instance_labels = virt.get_instance_labels(frag)
        instantiations = generate_instantiations(frag_circuit, instance_labels)
        num_instances += len(instantiations)
        jobs[frag] = virt.get_backend(frag).run(instantiations, shots=shots)
    logger.info(f"Running {num_instances} instances...")
    width = virt.num_clbits + len(virt._vgate_instrs)
    results = {}
    for frag, job in jobs.items():
        result = job.result()
        try:
            counts = result.get_counts()
        except Exception:                       # run.py:57-58: a fragment without measurements vanishes
            continue
        counts = [counts] if isinstance(counts, dict) else counts
        results[frag] = [QuasiDistr.from_counts(c, num_bits=width) for c in counts]
    run_time = perf_counter() - now
    logger.info("Knitting...")
    now = perf_counter()
    res_dist = virt.knit(results, None)
    knit_time = perf_counter() - now
    logger.info(f"Knitted in {knit_time:.2f}s.")
    return res_dist.nearest_probability_distribution(), RunTimeInfo(run_time, knit_time)
