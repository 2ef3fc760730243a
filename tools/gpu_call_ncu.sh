#!/bin/bash
# ncu --set full of the pair kernels of the current build (one launch each) -> gpurun_out/r1f_prof.ncu-rep
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on \
    -k regex:"tile_list_kernel|density_kernel|force_kernel" -s 9 -c 3 -f -o gpurun_out/r1f_prof \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r1f_ncu_full.log 2>&1
echo "ncu rc=$?"
ls -la gpurun_out/
tail -5 gpurun_out/r1f_ncu_full.log
