#!/bin/bash
# Last evidence pass of round 2: the knit kernel after the per-launch costs were removed (--set full + launch list).
tag=${1:-r2i}; out=gpurun_out/$tag; mkdir -p $out
timeout 300 ncu --set full --import-source on --clock-control none -k regex:knit_outer_kernel -c 1 -o $out/knit_outer python bench.py --workload syc32d1 --profile --steps 1 --warmup 3 > $out/ncu_knit_outer.log 2>&1
python tools/ncu_summary.py $out/knit_outer.ncu-rep > $out/knit_outer_summary.txt 2>&1
rm -f $out/knit_outer.ncu-rep
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_syc32d1.csv python bench.py --workload syc32d1 --profile --steps 2 --warmup 3 > $out/ncu_launches_syc32d1.log 2>&1
python tools/launch_summary.py $out/launches_syc32d1.csv > $out/launches_syc32d1_summary.txt 2>&1
cat $out/knit_outer_summary.txt $out/launches_syc32d1_summary.txt
