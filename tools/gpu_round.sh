#!/bin/bash
# One gpurun call: GPU tests, smoke, bench lines and ncu launch lists; everything lands in gpurun_out/<tag>/.
# usage: tools/gpu_round.sh <tag> [stages...]   stages: tests smoke bench all ref ncu
tag=${1:-r2}; shift
stages=${@:-tests smoke bench all ref ncu}
out=gpurun_out/$tag; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/gpu.txt 2>&1
for s in $stages; do
  t0=$(date +%s)
  case $s in
    tests) timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/summary.txt; tail -5 $out/pytest.log ;;
    testsall) timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/summary.txt; tail -15 $out/pytest.log ;;
    smoke) timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?" >> $out/summary.txt; tail -3 $out/smoke.log ;;
    bench) timeout 900 python bench.py > $out/bench_default.json 2> $out/bench_default.err; echo "bench rc=$?" >> $out/summary.txt ;;
    all) timeout 1200 python bench.py --workload all > $out/bench_all.json 2> $out/bench_all.err; echo "bench all rc=$?" >> $out/summary.txt ;;
    ref) timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err; echo "ref rc=$?" >> $out/summary.txt ;;
    ncu) for w in syc32d1 hwe16d5 syc16d5 bv16; do
           timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_$w.csv \
             python bench.py --workload $w --profile --steps 2 --warmup 3 > $out/ncu_$w.log 2>&1; echo "ncu $w rc=$?" >> $out/summary.txt
         done ;;
    *) echo "unknown stage $s" ;;
  esac
  echo "$s took $(( $(date +%s) - t0 )) s" >> $out/summary.txt
done
cat $out/summary.txt
