#!/bin/bash
# Second evidence pass of round 2: --set full captures of the kernels changed late in the round (knit_outer with the
# fused tail, npd_cluster_kernel) and of sim_tree_combine_kernel, with the per-instruction source page of the latter
# two.  Each command has already exited 0 without ncu.  usage: tools/ncu_round2.sh <tag>
tag=${1:-r2c}; out=gpurun_out/$tag; mkdir -p $out
full="ncu --set full --import-source on --clock-control none"
timeout 600 $full -k regex:knit_outer_kernel -c 1 -o $out/knit_outer python bench.py --workload syc32d1 --profile --steps 1 --warmup 3 > $out/ncu_knit_outer.log 2>&1
timeout 600 $full -k regex:npd_cluster -s 3 -c 1 -o $out/npd_cluster python bench.py --workload hwe16d5 --profile --steps 1 --warmup 3 > $out/ncu_npd_cluster.log 2>&1
timeout 600 $full -k regex:sim_tree_combine -s 6 -c 1 -o $out/tree_combine python bench.py --workload hwe16d5 --profile --steps 1 --warmup 3 > $out/ncu_tree_combine.log 2>&1
for r in knit_outer npd_cluster tree_combine; do
  python tools/ncu_summary.py $out/$r.ncu-rep > $out/${r}_summary.txt 2>&1
  ncu -i $out/$r.ncu-rep --page source --csv > $out/${r}_source.csv 2>/dev/null
done
for w in hwe16d5 bv16; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_$w.csv \
    python bench.py --workload $w --profile --steps 2 --warmup 3 > $out/ncu_launches_$w.log 2>&1
  python tools/launch_summary.py $out/launches_$w.csv > $out/launches_${w}_summary.txt 2>&1
done
rm -f $out/knit_outer.ncu-rep   # (large; the summary and the source page stay)
ls -la $out | tail -30
