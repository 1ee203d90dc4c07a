#!/bin/bash
# Evidence for profiles/: ncu launch lists (eager launches: --profile) and --set full captures of the dominant
# kernels.  Each command has already exited 0 without ncu (tools/gpu_round.sh).  usage: tools/ncu_round.sh <tag>
tag=${1:-r2}; out=gpurun_out/$tag; rep=/tmp/$tag; mkdir -p $out $rep   # .ncu-rep files stay on the box (64 MiB limit)
for w in syc32d1 hwe16d5 syc16d5 bv16; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_$w.csv \
    python bench.py --workload $w --profile --steps 2 --warmup 3 > $out/ncu_launches_$w.log 2>&1
  python tools/launch_summary.py $out/launches_$w.csv > $out/launches_${w}_summary.txt 2>&1
done
full="ncu --set full --import-source on --clock-control none"
timeout 600 $full -k regex:knit_outer_kernel -c 1 -o $rep/knit_outer python bench.py --workload syc32d1 --profile --steps 1 --warmup 3 > $out/ncu_knit_outer.log 2>&1
timeout 600 $full -k regex:"sim_tree_level|sim_tree_combine" -s 24 -c 12 -o $rep/sim_tree python bench.py --workload hwe16d5 --profile --steps 1 --warmup 3 > $out/ncu_sim_tree.log 2>&1
timeout 600 $full -k regex:contract_dmma_pipe -s 3 -c 1 -o $rep/contract_pipe python bench.py --workload hwe16d5 --profile --steps 1 --warmup 3 > $out/ncu_contract.log 2>&1
timeout 600 $full -k regex:"npd_" -s 24 -c 8 -o $rep/npd python bench.py --workload hwe16d5 --profile --steps 1 --warmup 3 > $out/ncu_npd.log 2>&1
QCK_SIM_TREE=0 timeout 600 $full -k regex:sim_warp -s 12 -c 4 -o $rep/sim_warp python bench.py --workload hwe16d5 --profile --steps 1 --warmup 3 > $out/ncu_sim_warp.log 2>&1
timeout 900 $full -k regex:"zero_dead|sim_sweep_tma" -c 4 -o $rep/uncut_sweeps python tools/prof_sweep.py syc 32 1 1 > $out/ncu_uncut.log 2>&1
for r in knit_outer sim_tree contract_pipe npd sim_warp uncut_sweeps; do
  python tools/ncu_summary.py $rep/$r.ncu-rep > $out/${r}_summary.txt 2>&1
done
ls -la $out $rep | tail -40
