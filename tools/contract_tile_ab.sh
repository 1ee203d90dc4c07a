#!/bin/bash
# A/B of the contraction tile size (QCK_CONTRACT_TILE) on the K >= 1 workloads + an ncu capture of the 128x128 kernel
tag=${1:-r2tile}; out=gpurun_out/$tag; mkdir -p $out
timeout 500 python -m pytest tests -m gpu -q -x --timeout 300 > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest.log
for w in hwe16d5 syc16d5 bv16 aqft16:solver; do
  for t in 64 128; do
    QCK_CONTRACT_TILE=$t timeout 300 python bench.py --workload $w --no-others --no-cpu-baseline --steps 20 --warmup 3 > $out/bench_${w}_$t.json 2> $out/bench_${w}_$t.err
    python - <<PY
import json
try:
    d=json.loads([x for x in open("$out/bench_${w}_$t.json") if x.startswith("{")][-1])
    print("$w tile $t:", round(d["ms_per_step"],4), "ms", {k:round(v,3) for k,v in d["roofline"]["share_of_step"].items()}, d.get("oracle",{}).get("max_abs_err_knit_vs_oracle"))
except Exception as e: print("$w $t failed", e)
PY
  done
done
ncu --set full --import-source on --clock-control none -k regex:contract_dmma_pipe -s 3 -c 1 -o $out/contract128 python bench.py --workload hwe16d5 --profile --steps 1 --warmup 3 > $out/ncu_contract128.log 2>&1
python tools/ncu_summary.py $out/contract128.ncu-rep > $out/contract128_summary.txt 2>&1
ncu -i $out/contract128.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
for name in h:
    if 'dmma' in name or 'lts__t_bytes.sum' == name or 'l1tex__m_xbar2l1tex_read_bytes.sum'==name:
        i=h.index(name); print(name, [r[i] for r in rows[1:]])
" > $out/contract128_dmma.txt 2>&1
cat $out/contract128_summary.txt $out/contract128_dmma.txt
rm -f $out/contract128.ncu-rep
