#!/bin/bash
# One gpurun call: variant sweep of the density/force passes -> GPU test suite and full bench line on the winner.
#   /usr/local/graft/bin/gpurun --timeout 840 -- 'bash tools/gpu_call_r1f.sh <variant names>'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r1f_gpu.txt 2>&1
timeout 420 python tools/variant_sweep.py run "$@" --budget=300 > gpurun_out/r1f_sweep.log 2>&1
W=$(cat gpurun_out/variant_winner.txt 2>/dev/null || echo base)
export PYTICLES_B200_LIB=$PWD/variants/libpyticles_b200_$W.so
echo "winner $W" | tee gpurun_out/r1f_winner.txt
timeout 330 python -m pytest tests -x -q -m gpu > gpurun_out/r1f_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r1f_winner.txt
tail -3 gpurun_out/r1f_pytest_gpu.log
timeout 200 python bench.py > gpurun_out/r1f_bench_c3.json 2> gpurun_out/r1f_bench_c3.err
echo "bench rc=$?" >> gpurun_out/r1f_winner.txt
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/r1f_launches_raw.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r1f_ncu_bench.log 2>&1
echo "ncu rc=$?" >> gpurun_out/r1f_winner.txt
cat gpurun_out/variant_sweep.txt
cat gpurun_out/r1f_bench_c3.json
