#!/bin/bash
tag=${1:-r2tile2}; out=gpurun_out/$tag; mkdir -p $out
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --workload hwe16d5 --no-others --no-cpu-baseline --steps 30 --warmup 3 > $out/bench_$name.json 2> $out/bench_$name.err
  python - <<PY
import json
try:
    d=json.loads([x for x in open("$out/bench_$name.json") if x.startswith("{")][-1])
    print("$name:", round(d["ms_per_step"],4), "ms", {k:round(v*d["ms_per_step"],4) for k,v in d["roofline"]["share_of_step"].items()})
except Exception as e: print("$name failed", e)
PY
}
run t64 QCK_CONTRACT_TILE=64
run t128w8 QCK_CONTRACT_TILE=128
run t128w16 QCK_CONTRACT_TILE=128 QCK_CONTRACT_WARPS=16
run t128w16s2 QCK_CONTRACT_TILE=128 QCK_CONTRACT_WARPS=16 QCK_CONTRACT_CTAS_PER_SM=1
run t64 QCK_CONTRACT_TILE=64
run t128w8 QCK_CONTRACT_TILE=128
run t128w16 QCK_CONTRACT_TILE=128 QCK_CONTRACT_WARPS=16
QCK_CONTRACT_WARPS=16 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:contract_dmma_pipe -s 3 -c 2 python bench.py --workload hwe16d5 --profile --steps 1 --warmup 3 2>&1 | grep -E "contract_dmma|duration|dmma" | head
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:contract_dmma_pipe -s 3 -c 2 python bench.py --workload hwe16d5 --profile --steps 1 --warmup 3 2>&1 | grep -E "contract_dmma|duration|dmma" | head
timeout 200 python -m pytest tests -m gpu -q -x --timeout 300 -k "contract or knit or baseline" 2>&1 | tail -3
QCK_CONTRACT_WARPS=16 timeout 200 python -m pytest tests -m gpu -q -x --timeout 300 -k "contract or knit or baseline" 2>&1 | tail -3
