#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_distributed.py -x -q -m gpu > gpurun_out/r1f_pytest_g2.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r1f_pytest_g2.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r1f_bench_g2.json 2> gpurun_out/r1f_bench_g2.err; echo "g2 rc=$?"
cut -c1-400 gpurun_out/r1f_bench_g2.json
