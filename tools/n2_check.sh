#!/bin/bash
# 2-GPU check: the GPU suite (incl. tests/test_gpu_multi.py at 2 ranks) and the bench line
mkdir -p gpurun_out/r2n2
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r2n2/pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2n2/pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2n2/bench_n2.json 2> gpurun_out/r2n2/bench_n2.err
echo "bench rc=$? $(tail -c 300 gpurun_out/r2n2/bench_n2.err | tr '\n' ' ')"
