#!/usr/bin/env python
"""bench.py - BASELINE.json metric: syc-32 d1 fragment simulation + knit wall time.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload syc32d1] [--impl ours|reference]

One "step" = one pass of the hot path over the workload: simulate every fragment
instance, knit the fragment tables into the dense full-circuit distribution and
reduce its (sum, min).  Default workload = ``syc32d1`` (``benchmark.py -p 2 -q 50
syc 32 1``, seeded): 18- and 14-qubit fragments, no virtual gates, a 2^32-entry
(32 GiB) float64 result - it fits one B200.  Under torchrun (N > 1) the output
index is sharded by its top bits (total work fixed -> "strong" scaling); the
16-qubit configs shard the label range instead and all-reduce the result.

JSON line (rank 0): ``value`` = seconds per step with programs resident in HBM
(CUDA events, max over ranks); ``e2e`` = the same through the public API
``run_virtual_circuit_dense`` on a fresh ``VirtualCircuit`` each step (host
compile, pinned H2D of programs, kernels, D2H of the statistics);
``roofline`` = the knit kernel's algorithmic bytes / its own CUDA-event time
against the measured HBM copy peak (MEASURED_PEAKS.json); ``cpu_baseline`` = the
oracle port (C + OpenMP, all host cores) on a bounded sample, extrapolated.

``--impl reference`` times that CPU path alone (the reference's own CPU
implementation - qiskit-aer + multiprocessing - cannot be installed in this
image) and prints the same line with ``"impl": "reference"``.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "syc-32 d1 sim+knit wall time (s) at 1/2/4/8 B200; HBM GB/s; fidelity delta vs ref"
PKG = "hardwareawareoptimalquantumcircuitcuttingandknitting_b200"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="syc32d1",
                    help="a BASELINE config (cutting.BASELINE_CONFIGS); 'NAME:solver' cuts it with the z3 cutter "
                         "instead of applying the recorded cut shape (e.g. aqft16:solver: five wire cuts)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-sample-bits", type=int, default=None,
                    help="log2 of the output entries the CPU baseline knits per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--uncut-statevector", action="store_true", default=None,
                    help="also simulate the UNCUT circuit as one statevector on this GPU (64 GiB at 32 qubits), "
                         "report the streaming simulator against the HBM roofline and the dense fidelity "
                         "(default: on for one GPU outside --profile)")
    ap.add_argument("--no-uncut-statevector", dest="uncut_statevector", action="store_false")
    ap.add_argument("--profile", action="store_true",
                    help="timed region only (for runs under ncu): no e2e leg, no oracle report, no CPU baseline")
    return ap.parse_args()


def metric_name(workload: str) -> str:
    return METRIC if workload == "syc32d1" else f"{workload} sim+knit wall time (s)"


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region: an NVML polling thread (a few ms per sample;
    the nvidia-smi process of the profiling recipe needs longer to start than a 100 ms timed region lasts)."""
    REASONS = (("hw_slowdown", "HwSlowdown"), ("hw_thermal_slowdown", "HwThermalSlowdown"),
               ("sw_thermal_slowdown", "SwThermalSlowdown"), ("sw_power_cap", "SwPowerCap"),
               ("hw_power_brake_slowdown", "HwPowerBrakeSlowdown"))

    def __init__(self, gpu_index: int, period_ms: int = 2) -> None:
        import threading
        self.rows, self.err, self.nv, self.handle = [], None, None, None
        self._stop = threading.Event()
        self.period = period_ms / 1e3
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = gpu_index
            if visible:
                try:
                    phys = int(visible.split(",")[gpu_index])
                except (ValueError, IndexError):
                    phys = gpu_index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        except Exception as exc:      # no NVML: say so in the bench line
            self.err = repr(exc)

    def _sample(self) -> None:
        nv = self.nv
        mhz = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
        try:
            mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
        except Exception:
            mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        self.rows.append((time.time(), mhz, mask))

    def _run(self) -> None:
        while not self._stop.is_set():
            try:
                self._sample()
            except Exception as exc:
                self.err = repr(exc)
                return
            self._stop.wait(self.period)

    def stop(self, t0: float, t1: float) -> dict:
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"NVML unavailable: {self.err}"]}
        self._stop.set()
        self.thread.join(timeout=2)
        in_region = [r for r in self.rows if t0 <= r[0] <= t1]
        use = in_region or self.rows
        if not use:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [f"no samples ({self.err})"]}
        reasons = set()
        for _, _, mask in use:
            for name, suffix in self.REASONS:
                bit = getattr(self.nv, "nvmlClocksEventReason" + suffix, None)
                if bit is None:
                    bit = getattr(self.nv, "nvmlClocksThrottleReason" + suffix, 0)
                if mask & bit:
                    reasons.add(name)
        return {"sm_mhz": statistics.median(r[1] for r in use), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(reasons), "samples": len(use), "samples_in_timed_region": len(in_region),
                "source": "NVML polling thread"}


# ----------------------------------------------------------------------------- CPU baseline (oracle port)
def _fragment_out_mask(ov, frag) -> int:
    """clbits the fragment's own (final) measurements write (bit i = clbit i)."""
    cidx = {c: i for i, c in enumerate(ov.circuit.clbits)}
    mask = 0
    for op in ov.frag_ops[frag]:
        if op.operation is not None and getattr(op.operation, "name", "") == "measure":
            mask |= 1 << cidx[op.clbits[0]]
    return mask


def cpu_fragment_table_k0(ov, frag):
    """Exact table of a fragment without virtual gates: C oracle statevector, |amp|^2 compacted
    to the measured clbits in ascending order.  -> (table, clbit mask)"""
    import numpy as np
    from oracle import cport
    inst = ov.instance(frag, ())
    qidx = {q: i for i, q in enumerate(inst.qubits)}
    cidx = {c: i for i, c in enumerate(inst.clbits)}
    last = {}
    for i, d in enumerate(inst.data):
        for q in d.qubits:
            last[q] = i
    pairs = []
    for i, d in enumerate(inst.data):
        if getattr(d.operation, "name", "") == "measure":
            assert last[d.qubits[0]] == i, "C-oracle fast path needs terminal measurements"
            pairs.append((cidx[d.clbits[0]], qidx[d.qubits[0]]))
    pairs.sort()
    mask = 0
    for c, _ in pairs:
        mask |= 1 << c
    prob = cport.simulate_probabilities(inst)
    if [q for _, q in pairs] == list(range(len(inst.qubits))):
        return prob, mask
    idx = np.arange(1 << len(inst.qubits), dtype=np.uint64)
    comp = np.zeros_like(idx)
    for j, (_, q) in enumerate(pairs):
        comp |= ((idx >> np.uint64(q)) & np.uint64(1)) << np.uint64(j)
    return np.bincount(comp.astype(np.int64), weights=prob, minlength=1 << len(pairs)), mask


def cpu_reference_step(workload: str, seed: int, sample_bits: int | None):
    """One bounded CPU step of the workload with the oracle port.  Returns
    (extrapolated seconds for the whole job, description of the sample, cores)."""
    import numpy as np
    from math import cos, sin
    from importlib import import_module
    from oracle import cport, dense as od, instantiate as oi, qpd_tables as qt, statevector as sv
    cutting = import_module(f"{PKG}.cutting")
    cores = cport.num_threads()
    circ, cut = cutting.make_baseline(workload, seed)
    ov = oi.OracleVirtualCircuit(cut)
    frags = [f for f in ov.fragments if any(ov.has_measurement(f, l) for l in ov.instance_labels(f))]
    K = len(ov.vgates)
    n_cl = ov.n_clbits
    t0 = time.perf_counter()
    if K == 0:
        pairs = [cpu_fragment_table_k0(ov, f) for f in frags]
        tabs, masks = [p[0] for p in pairs], [p[1] for p in pairs]
        t_sim = time.perf_counter() - t0
        n_out = bin(sum(masks)).count("1")
        sb = min(n_out, sample_bits if sample_bits is not None else 26)
        t1 = time.perf_counter()
        cport.knit_outer(tabs, masks, 0, 1 << sb, want_output=True)
        t_knit = time.perf_counter() - t1
        scale = float(1 << (n_out - sb))
        total = t_sim + t_knit * scale
        sample = (f"full simulation of {len(frags)} fragments ({t_sim:.3f} s) + outer-product knit of the first "
                  f"2^{sb} of 2^{n_out} output entries ({t_knit:.3f} s), knit time scaled x{scale:.0f}")
        return total, sample, cores
    # virtual gates: a time-bounded number of instances per fragment, then the full contraction
    folded, masks, touches = [], [], []
    t_sim, n_done, n_total = 0.0, 0, 0
    budget_s = 6.0
    for frag in frags:
        labels = ov.instance_labels(frag)
        n_total += len(labels)
        mask = _fragment_out_mask(ov, frag)
        full = np.zeros((len(labels), 1 << bin(mask).count("1")))
        spent = 0.0
        for i, lab in enumerate(labels):
            if spent > budget_s:
                break
            ts = time.perf_counter()
            full[i] = od.signed_fold(sv.exact_distribution(ov.instance(frag, lab)), n_cl, K, mask)
            spent += time.perf_counter() - ts
            n_done += 1
        t_sim += spent
        folded.append(full)
        masks.append(mask)
        touches.append(ov.touches(frag))
    sim_scale = n_total / max(n_done, 1)
    radices = ov.radices
    coeffs = []
    for (kind, theta, _), r in zip(ov.vgates, radices):
        if kind in ("rzz", "cp"):
            m = qt.knit_param(kind, theta)
            c, s_ = cos(m / 2), sin(m / 2)
            coeffs.append([c * c, s_ * s_, c * s_, c * s_, -c * s_, -c * s_][:r])
        else:
            coeffs.append([0.5 * sg for sg in (1, 1, 1, -1, 1, -1, 1, -1)[:r]])
    L = int(np.prod(radices))
    digits = np.stack(np.unravel_index(np.arange(L), radices), axis=1)
    w = np.ones(L)
    for k in range(K):
        w *= np.asarray(coeffs[k])[digits[:, k]]
    rows_idx = []
    for touch in touches:
        st, acc = [0] * K, 1
        for k in reversed(range(K)):
            if touch[k]:
                st[k] = acc
                acc *= radices[k]
        rows_idx.append(digits @ np.asarray(st))
    n_out = bin(sum(masks)).count("1")
    t1 = time.perf_counter()
    cport.knit_contract(folded, masks, n_out, w, np.stack(rows_idx))
    t_knit = time.perf_counter() - t1
    total = t_sim * sim_scale + t_knit
    sample = (f"{n_done} of {n_total} fragment instances simulated (numpy oracle, {t_sim:.2f} s, scaled "
              f"x{sim_scale:.1f}) + full dense contraction over {L} labels (C oracle, {t_knit:.2f} s)")
    return total, sample, cores


def reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    times, sample, cores = [], "", 1
    for i in range(args.warmup + args.steps):
        t, sample, cores = cpu_reference_step(args.workload, args.seed, args.cpu_sample_bits)
        if i >= args.warmup:
            times.append(t)
        if sum(times) > 240:
            break
    val = statistics.mean(times)
    line = {
        "impl": "reference", "metric": metric_name(args.workload), "value": val, "unit": "s",
        "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup, "ms_per_step": val * 1e3,
        "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": args.workload, "seed": args.seed},
        "cpu_baseline": {"value": val, "unit": "s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference CPU path (qiskit-aer + multiprocessing) is not installable here; this is the "
                "oracle port (C/OpenMP + numpy) on all host cores",
    }
    emit(line)


# ----------------------------------------------------------------------------- our arm
_REAL_STDOUT = None


def _quiet_stdout() -> None:
    """Libraries (NCCL's version banner, torchrun) write to fd 1; the contract is ONE JSON line on
    stdout.  Point fd 1 at stderr for the whole run and keep the real stdout for the final line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main() -> None:
    args = parse_args()
    _quiet_stdout()
    if args.impl == "reference":
        reference_arm(args)
        return
    import numpy as np
    import torch
    from importlib import import_module
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback")
    qdist = import_module(f"{PKG}.dist")
    cutting = import_module(f"{PKG}.cutting")
    vc = import_module(f"{PKG}.virtual_circuit")
    runm = import_module(f"{PKG}.run")
    fid = import_module(f"{PKG}.fidelity")
    lib = import_module(f"{PKG}._lib")
    import torch.distributed as dist

    rank, local_rank, world = qdist.init_from_env("nccl")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    handle = lib.get_handle(local_rank)

    circ, cut = cutting.make_baseline(args.workload, args.seed)
    virt = vc.VirtualCircuit(cut)
    K = len(virt.vgates)
    masks, union = virt.output_masks()
    n_out = bin(union).count("1")
    frags = virt.active_fragments()
    L = virt.num_global_labels()

    if K == 0:
        y0, y1 = qdist.shard_pow2(n_out, rank, world) if world > 1 else (0, 1 << n_out)
        label_range = None
        out = torch.empty(y1 - y0, dtype=torch.float64, device=device)
    else:
        y0, y1 = 0, 1 << n_out
        label_range = qdist.shard_range(L, rank, world, align=virt.global_radices()[-1]) if world > 1 else None
        out = torch.empty(1 << n_out, dtype=torch.float64, device=device)
    stats = torch.zeros(4, dtype=torch.float64, device=device)
    # programs resident in HBM before the timed region
    for f in frags:
        virt.executor(f, device, True).upload()
    tables_holder = {}
    knit_events = []

    def step_resident(record: bool) -> None:
        tables = virt.simulate_fragments(device, label_range=label_range)
        tables_holder["t"] = tables
        if record:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        if K == 0:
            virt.knit_tables(tables, device, stats=stats, y_range=(y0, y1) if world > 1 else None, out=out)
        else:
            virt.knit_tables(tables, device, label_range=label_range, out=out,
                             stats=None if world > 1 else stats)
        if record:
            e1.record()
            knit_events.append((e0, e1))
        if world > 1:
            if K == 0:
                qdist.allreduce_stats(stats)
            else:
                qdist.allreduce_sum_(out)

    def barrier() -> None:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    for _ in range(max(args.warmup, 3)):
        step_resident(False)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    time.sleep(0.15 if sampler else 0.0)
    barrier()
    launches0 = handle.launch_count
    t_wall0 = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step_resident(True)
    ev1.record()
    barrier()
    t_wall1 = time.time()
    launches = handle.launch_count - launches0
    elapsed_ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=device)
    knit_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in knit_events) / len(knit_events)],
                           dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(elapsed_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(knit_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(elapsed_ms.item()) / args.steps
    knit_ms = float(knit_ms.item())
    host_stats = stats.cpu().numpy().copy()
    if K > 0 and world > 1:
        handle.check(handle.lib.qck_stats_dense(handle.ptr, out.data_ptr(), out.numel(), 0.0, stats.data_ptr(),
                                                torch.cuda.current_stream(device).cuda_stream))
        host_stats = stats.cpu().numpy().copy()

    # ---- e2e: public API, fresh VirtualCircuit per step (host compile + H2D + kernels + D2H)
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    e2e_s, e2e_warm_s, h2d = None, None, 0
    if not args.profile:
        e2e_virts = [vc.VirtualCircuit(cut) for _ in range(args.steps + 1)]
        runm.run_virtual_circuit_dense(e2e_virts[0], device=device, rank=rank, world_size=world, out=out,
                                       nearest=False)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for v in e2e_virts[1:]:
            vc.clear_program_cache()            # e2e is the COLD path: every step compiles its programs
            runm.run_virtual_circuit_dense(v, device=device, rank=rank, world_size=world, out=out, nearest=False)
        g1.record()
        barrier()
        # same call with the process-wide program cache warm (what a second run of the same cut
        # circuit costs, e.g. the reference's ideal + noisy pair)
        warm_virts = [vc.VirtualCircuit(cut) for _ in range(args.steps)]
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record()
        for v in warm_virts:
            runm.run_virtual_circuit_dense(v, device=device, rank=rank, world_size=world, out=out, nearest=False)
        w1.record()
        barrier()
        warm_ms = torch.tensor([w0.elapsed_time(w1)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(warm_ms, op=dist.ReduceOp.MAX)
        e2e_warm_s = float(warm_ms.item()) / args.steps / 1e3
        for f in e2e_virts[1].active_fragments():
            h2d += e2e_virts[1].executor(f, device, True).h2d_bytes
        e2e_ms = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
        e2e_s = float(e2e_ms.item()) / args.steps / 1e3

    # ---- fidelity against the uncut circuit and parity against the oracle (outside the timed region)
    extra = {}
    if rank == 0 and not args.profile:
        extra = gpu_fidelity_report(virt, circ, tables_holder["t"], out, device, fid, vc, handle, K, n_out, world)

    if args.uncut_statevector is None:
        args.uncut_statevector = not args.profile
    if rank == 0 and world == 1 and args.uncut_statevector:
        extra.update(uncut_statevector_report(circ, out, device, fid, vc, handle, peaks_hbm()))

    if rank != 0:
        if world > 1:
            dist.barrier()
        return
    # ---- roofline of the dominant kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    if K == 0:
        table_bytes = sum(8 * (1 << bin(m).count("1")) for m in masks.values())
        alg_bytes = 8 * (y1 - y0) + table_bytes
        kernel = "knit_outer_kernel"
    else:
        alg_bytes = 8 * (1 << n_out) + sum(int(t.numel()) * 8 for t in tables_holder["t"].values())
        kernel = "contract_gemm_kernel+contract_scatter_kernel"
    achieved = alg_bytes / (knit_ms * 1e-3) / 1e9
    traffic = None
    try:        # per-launch DRAM bytes of the same kernel from the committed ncu --set full capture
        if world == 1:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[args.workload][kernel]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": knit_ms,
                "kernel_share_of_step": knit_ms / ms_per_step}

    # ---- CPU-baseline leg (rank 0, N = 1): the oracle port timed on the host cores, and the parity
    #      of this run's GPU results against the oracle.  The only place oracle/ is touched.
    cpu = None
    if world == 1 and not args.no_cpu_baseline and not args.profile:
        try:
            val, sample, cores = cpu_reference_step(args.workload, args.seed, args.cpu_sample_bits)
            cpu = {"value": val, "unit": "s", "cores": cores, "kind": "port", "sample": sample}
        except Exception as exc:  # the baseline must never take the bench line down
            cpu = {"value": None, "unit": "s", "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {exc}"}
        extra.update(oracle_parity_report(virt, circ, cut, tables_holder["t"], out, y0, y1, device, K, n_out,
                                          extra.get("fidelity_cut_vs_uncut")))

    line = {
        "metric": metric_name(args.workload), "value": ms_per_step / 1e3, "unit": "s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
        "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": args.workload, "seed": args.seed, "n_out_bits": n_out,
                   "fragments": [len(f) for f in frags], "virtual_gates": K, "global_labels": L,
                   "partition": ("output index by top bits" if K == 0 else "label range + all-reduce"),
                   "l2": (f"every step rewrites the {8 * (y1 - y0) / 2**30:.2f} GiB result (>> 126 MB L2): "
                          "no flush needed" if 8 * (y1 - y0) > (1 << 28) else
                          "working set fits L2; small config, launch/latency bound")},
        "clocks": clocks,
        "e2e": {"value": e2e_s, "unit": "s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 32,
                "program_cache": "cold (cleared every step)", "value_program_cache_warm": e2e_warm_s},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "hbm_gbs": achieved,
        "result_sum": float(host_stats[0]), "result_min": float(host_stats[1]),
    }
    line.update(extra)
    emit(line)
    if world > 1:
        dist.barrier()


def peaks_hbm() -> float:
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def uncut_statevector_report(circ, cut_result, device, fid, vc, handle, peak) -> dict:
    """the uncut circuit as ONE statevector on this GPU (Utilities.py:39-69 runs it on
    Aer).  Streaming regime; the sweeps run on the TMA kernel with live-qubit tracking when eligible.  The
    bytes are the exact HBM traffic of the sweeps (qck_sim_plan_traffic), not a full read + write per
    sweep."""
    import ctypes as C
    import torch
    rep = {}
    try:
        virt_u = vc.VirtualCircuit(circ)                       # one register = one fragment, no cuts
        (frag,) = virt_u.active_fragments()
        ex = virt_u.executor(frag, device, True)
        n = ex.max_state
        plan = ex.plans[0]
        sweeps = len(plan.sweeps)
        ex.upload()
        table = ex.run(handle)                                 # warm-up (allocates state + row)
        torch.cuda.synchronize(device)
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        reps = 3
        e0.record()
        for _ in range(reps):
            ex.run(handle, out=table)
        e1.record()
        info = {"qubits": n, "sweeps": sweeps, "ops": int(len(plan.ops))}
        if ex.streaming:
            st = ex.plan_struct(0)
            stream = torch.cuda.current_stream(device).cuda_stream
            for _ in range(reps):                              # the sweeps alone, into the same state buffer
                handle.check(handle.lib.qck_sim_statevector(handle.ptr, C.byref(st), 0, ex._work.data_ptr(),
                                                            ex._work.numel(), stream))
            e2.record()
            torch.cuda.synchronize(device)
            ms, ms_sweeps = e0.elapsed_time(e1) / reps, e1.elapsed_time(e2) / reps
            ld, sd, used = C.c_uint64(), C.c_uint64(), C.c_int()
            handle.check(handle.lib.qck_sim_plan_traffic(C.byref(st), 1, 0, C.byref(ld), C.byref(sd), C.byref(used)))
            moved = ld.value + sd.value
            # the whole run: when the output row is the whole register the last sweep stores probabilities
            # (fold fused); else a separate pass reads the state and writes the row
            fused = bool(used.value) and not ex.program.radix and list(plan.out_pos) == list(range(n)) \
                and plan.sum_mask == 0 and os.environ.get("QCK_FOLD_FUSION", "1") != "0"
            ld2, sd2 = C.c_uint64(), C.c_uint64()
            handle.check(handle.lib.qck_sim_plan_traffic(C.byref(st), 1, int(fused), C.byref(ld2), C.byref(sd2), C.byref(used)))
            total_bytes = ld2.value + sd2.value + (0 if fused else (16 << n) + (8 << len(plan.out_pos)))
            info.update({"ms": ms, "bytes": total_bytes, "gbs": total_bytes / ms / 1e6,
                         "frac_of_measured_hbm_peak": total_bytes / ms / 1e6 / peak,
                         "fold": "fused into the last sweep (probabilities stored instead of amplitudes)" if fused
                         else "separate pass (reads the state, writes the row)",
                         "sweep_kernel": "sim_sweep_tma_kernel" if used.value else "sim_sweep_kernel",
                         "live_qubit_tracking": bool(used.value),
                         "statevector_sweeps": {"ms": ms_sweeps, "bytes_loaded": ld.value, "bytes_stored": sd.value,
                                                "gbs": moved / ms_sweeps / 1e6,
                                                "frac_of_measured_hbm_peak": moved / ms_sweeps / 1e6 / peak,
                                                "note": "qck_sim_statevector: the same sweeps leaving the amplitudes"},
                         "full_sweep_traffic_bytes": (16 << n) * (2 * sweeps - 1)})
        else:
            torch.cuda.synchronize(device)
            info["ms"] = e0.elapsed_time(e1) / reps
        rep["uncut_statevector"] = info
        if cut_result.numel() == table.numel():
            rep["fidelity_cut_vs_uncut_dense_statevector"] = fid.hellinger_fidelity(cut_result, table[0])
            worst, step = 0.0, 1 << 28                         # chunked: no 32 GiB temporary
            for a in range(0, table.numel(), step):
                worst = max(worst, float((cut_result[a:a + step] - table[0][a:a + step]).abs().max().item()))
            rep["max_abs_err_cut_vs_uncut_dense"] = worst
    except Exception as exc:
        rep["uncut_statevector_error"] = repr(exc)
    return rep


def gpu_fidelity_report(virt, circ, tables, out, device, fid, vc, handle, K, n_out, world) -> dict:
    """Hellinger fidelity of the cut result to the UNCUT circuit, both computed on the GPU by the
    product path only (Utilities.py:224 compares the ideal uncut run with the knitted one)."""
    import torch
    rep = {}
    try:
        masks, union = virt.output_masks()
        frags = list(tables.keys())
        if K == 0:
            # uncut circuit simulated per connected component, compared in factorised form
            comp_tabs, comp_masks = uncut_component_tables(circ, device)
            sp, sq, bc = fid.hellinger_fidelity_factored([tables[f][0] for f in frags], [masks[f] for f in frags],
                                                         comp_tabs, comp_masks, n_out, device)
            rep["fidelity_cut_vs_uncut"] = (bc / (sp * sq) ** 0.5) ** 2
        else:
            uncut_virt = vc.VirtualCircuit(circ)
            q = uncut_virt.knit_tables(uncut_virt.simulate_fragments(device), device)
            p = out.clone()
            stream = torch.cuda.current_stream(device).cuda_stream
            handle.check(handle.lib.qck_npd(handle.ptr, p.data_ptr(), p.numel(), 0.0, None, None, stream))
            rep["fidelity_cut_vs_uncut"] = fid.hellinger_fidelity(p, q)
            rep["max_abs_err_cut_vs_uncut_gpu"] = float((out - q).abs().max().item())
    except Exception as exc:
        rep["fidelity_report_error"] = repr(exc)
    return rep


def oracle_parity_report(virt, circ, cut, tables, out, y0, y1, device, K, n_out, f_gpu) -> dict:
    """Part of the CPU-baseline leg (the only place bench.py touches oracle/): the GPU results of
    this run against the oracle, on what the oracle finishes in seconds."""
    import numpy as np
    from importlib import import_module
    from oracle import cport, dense as od, instantiate as oi, statevector as sv
    rep = {}
    try:
        masks, union = virt.output_masks()
        frags = list(tables.keys())
        if K == 0:
            ov = oi.OracleVirtualCircuit(cut)
            pairs = [cpu_fragment_table_k0(ov, f) for f in ov.fragments
                     if ov.has_measurement(f, ())]
            o_tabs, o_masks = [p[0] for p in pairs], [p[1] for p in pairs]
            worst = 0.0
            by_mask = {masks[f]: tables[f][0].cpu().numpy() for f in frags}
            for t, m in zip(o_tabs, o_masks):
                worst = max(worst, float(np.abs(by_mask[m] - t).max()))
            rep["max_abs_err_fragment_tables_vs_oracle"] = worst
            # knit: a window of the output vs the oracle's outer product of ITS OWN tables
            win = min(1 << 20, y1 - y0)
            ref, _, _ = cport.knit_outer(o_tabs, o_masks, y0, y0 + win)
            got = out[:win].cpu().numpy()
            rep["max_abs_err_knit_window_vs_oracle"] = float(np.abs(got - ref).max())
            rep["knit_window"] = [int(y0), int(y0 + win)]
            # oracle fidelity, independent of every GPU table: both sides factorise over the connected
            # components of the uncut circuit, so BC = prod_c sum_x sqrt(p_c(x) q_c(x)) with q_c the
            # oracle's own simulation of component c and p_c the marginal of the oracle's fragment table
            cutting = import_module(f"{PKG}.cutting")
            ovc = oi.OracleVirtualCircuit(cutting.apply_cuts(circ, cutting.CutSpec()))
            f_or = 1.0
            for cfrag in ovc.fragments:
                if not ovc.has_measurement(cfrag, ()):
                    continue
                q_c, m_c = cpu_fragment_table_k0(ovc, cfrag)
                (t_f, m_f), = [(t, m) for t, m in zip(o_tabs, o_masks) if m & m_c]
                assert (m_c & ~m_f) == 0
                local = int(od.pext(np.uint64(m_c), m_f))
                idx = od.pext(np.arange(len(t_f), dtype=np.uint64), local).astype(np.int64)
                p_c = np.bincount(idx, weights=t_f, minlength=len(q_c))
                f_or *= float(np.sum(np.sqrt(p_c * q_c)) / np.sqrt(p_c.sum() * q_c.sum()))
            rep["fidelity_oracle"] = f_or ** 2
        else:
            uncut = sv.dense(sv.exact_distribution(circ), circ.num_clbits)
            got = out.cpu().numpy()
            rep["max_abs_err_vs_uncut_oracle"] = float(np.abs(got - uncut).max())
            rep["fidelity_oracle"] = od.hellinger_fidelity_dense(od.nearest_probability_distribution(got), uncut)
        if f_gpu is not None:
            rep["fidelity_delta_vs_oracle"] = abs(f_gpu - rep["fidelity_oracle"])
    except Exception as exc:
        rep["oracle_parity_error"] = repr(exc)
    return rep


def uncut_component_tables(circ, device, max_bits: int = 16, max_groups: int = 6):
    """Uncut circuit simulated on the GPU one connected component at a time (SURVEY.md 8f-1:
    until the sharded 32-qubit statevector exists), then multiplied into <= max_groups
    product tables of <= max_bits bits each (qck_knit_outer takes at most 8 tables)."""
    import ctypes as C
    import torch
    from importlib import import_module
    cutting = import_module(f"{PKG}.cutting")
    vc = import_module(f"{PKG}.virtual_circuit")
    lib = import_module(f"{PKG}._lib")
    comp = cutting.apply_cuts(circ, cutting.CutSpec())          # no cuts: one register per component
    v = vc.VirtualCircuit(comp)
    tabs = v.simulate_fragments(device)
    masks, _ = v.output_masks()
    items = sorted(((masks[f], tabs[f][0]) for f in tabs), key=lambda t: t[0] & -t[0])
    groups = []                                                  # [mask, [(mask, table)]]
    for m, t in items:
        for g in groups:
            if bin(g[0] | m).count("1") <= max_bits and len(g[1]) < 8:
                g[0] |= m
                g[1].append((m, t))
                break
        else:
            groups.append([m, [(m, t)]])
    if len(groups) > max_groups:
        raise NotImplementedError("uncut circuit has too many / too wide components for the factorised check")
    handle = lib.get_handle(device.index or 0)
    stream = torch.cuda.current_stream(device).cuda_stream
    out_tabs, out_masks = [], []
    for gmask, members in groups:
        nb = bin(gmask).count("1")
        ptrs = (C.c_void_p * len(members))(*[t.data_ptr() for _, t in members])
        cm = (C.c_uint64 * len(members))(*[vc._compress_mask(m, gmask) for m, _ in members])
        table = torch.empty(1 << nb, dtype=torch.float64, device=device)
        handle.check(handle.lib.qck_knit_outer(handle.ptr, len(members), ptrs, cm, nb, 0, 1 << nb,
                                               table.data_ptr(), None, stream))
        out_tabs.append(table)
        out_masks.append(gmask)
    return out_tabs, out_masks


if __name__ == "__main__":
    main()
