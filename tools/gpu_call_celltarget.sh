#!/bin/bash
mkdir -p gpurun_out
for t in 0 8; do
  SPH_CELL_TARGET=$t timeout 60 python tools/nl_sweep.py --logn 22 --reps 5 > gpurun_out/celltarget_sweep_$t.txt 2>&1
  SPH_CELL_TARGET=$t timeout 40 python bench.py --workload c2 --steps 10 --no-e2e --no-cpu-baseline > gpurun_out/celltarget_c2_$t.json 2> gpurun_out/celltarget_c2_$t.err
done
SPH_CELL_TARGET=8 timeout 60 python -m pytest tests -x -q -m gpu > gpurun_out/celltarget_pytest.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/celltarget_pytest.log
for t in 0 8; do echo "target $t"; grep -v "^#" gpurun_out/celltarget_sweep_$t.txt; python -c "
import json; js=json.load(open('gpurun_out/celltarget_c2_$t.json')); print('c2', js['ms_per_step'], {k:v['ms'] for k,v in js['roofline']['passes'].items()})"; done
