#!/bin/bash
# One gpurun call (one GPU): GPU suite, smoke, the bench line with every config, the launch list and the `ncu --set full`
# capture of the pair kernels of the same build -> gpurun_out/r2_*; profiles/summarise.py turns them into profiles/r2_*.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_call_r2.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2_gpu.txt 2>&1
timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/r2_pytest_gpu.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2_smoke.log
timeout 400 python bench.py --all-configs > gpurun_out/r2_bench_c3.json 2> gpurun_out/r2_bench_c3.err; echo "bench rc=$?"
timeout 100 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference.json 2>/dev/null; echo "reference rc=$?"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/r2_launches_raw.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-check > gpurun_out/r2_ncu_bench.log 2>&1; echo "launch list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"tile_list_kernel|density_kernel|force_kernel" \
    -s 9 -c 3 -f -o gpurun_out/r2_prof python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-check > gpurun_out/r2_ncu_full.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/r2_prof.ncu-rep
