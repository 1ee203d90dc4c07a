#!/bin/bash
mkdir -p gpurun_out/r2n2
echo skip-pytest
for i in 1 2 3 4; do
  QCK_BENCH_FAITHFUL_MULTI=1 QCK_BENCH_STAGE_SYNC=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500+i)) bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2n2/bench_n2_nosync_$i.json 2> gpurun_out/r2n2/bench_n2_nosync_$i.err
  echo "run $i rc=$? $(tail -c 300 gpurun_out/r2n2/bench_n2_nosync_$i.err | tr '\n' ' ')"
done
