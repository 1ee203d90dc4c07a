#!/bin/bash
# bench.py at N GPUs of one box (the driver's launch line), plus the reference arm; usage: tools/scale_round.sh <N> <tag>
n=${1:-2}; tag=${2:-r2scale}; out=gpurun_out/$tag; mkdir -p $out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $n --steps 20 --warmup 3 > $out/bench_n$n.json 2> $out/bench_n$n.err
echo "bench n=$n rc=$? $(tail -c 200 $out/bench_n$n.err | tr '\n' ' ')"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus $n --steps 2 --warmup 1 > $out/bench_reference_n$n.json 2> $out/bench_reference_n$n.err
echo "reference n=$n rc=$?"
