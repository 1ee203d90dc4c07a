#!/bin/bash
out=gpurun_out/r2fuse; mkdir -p $out
timeout 400 python -m pytest tests -m gpu -q -x --timeout 300 > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $out/pytest.log
for w in hwe16d5 syc16d5 bv16 aqft16:solver; do
  for f in 0 300 default; do
    if [ $f = default ]; then unset QCK_TREE_FUSE_ITEMS; else export QCK_TREE_FUSE_ITEMS=$f; fi
    timeout 300 python bench.py --workload $w --no-others --no-cpu-baseline --steps 30 --warmup 3 > $out/bench_${w}_$f.json 2> $out/bench_${w}_$f.err
    python - <<PY
import json
try:
    d=json.loads([x for x in open("$out/bench_${w}_$f.json") if x.startswith("{")][-1])
    print("$w fuse $f:", round(d["ms_per_step"],4), "ms", {k:round(v*d["ms_per_step"],4) for k,v in d["roofline"]["share_of_step"].items()}, "launches", d["gpu_launches"], d.get("oracle",{}).get("max_abs_err_knit_vs_oracle"))
except Exception as e: print("$w $f failed", e)
PY
  done
done
