#!/bin/bash
mkdir -p gpurun_out
timeout 60 python -m pytest tests -x -q -m gpu > gpurun_out/r1f_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -1 gpurun_out/r1f_pytest_gpu.log
timeout 40 python bench.py --workload c2 --no-cpu-baseline > gpurun_out/r1f_bench_c2.json 2> gpurun_out/r1f_bench_c2.err; echo "c2 rc=$?"
cut -c1-200 gpurun_out/r1f_bench_c2.json
