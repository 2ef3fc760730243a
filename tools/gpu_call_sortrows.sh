#!/bin/bash
mkdir -p gpurun_out
for v in 0 1; do
  SPH_SORT_ROWS=$v timeout 200 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/sortrows_$v.json 2> gpurun_out/sortrows_$v.err
  python - <<PY
import json
js=json.load(open("gpurun_out/sortrows_$v.json"))
print("SPH_SORT_ROWS=$v", js["ms_per_step"], {k:v["ms"] for k,v in js["roofline"]["passes"].items()})
PY
done
