#!/bin/bash
# Final single-GPU check of the current build: GPU test suite, smoke, bench lines (c3 full, c2, c4, reference arm).
mkdir -p gpurun_out
timeout 200 python -m pytest tests -x -q -m gpu > gpurun_out/r1f_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/r1f_pytest_gpu.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1f_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r1f_smoke.log
timeout 200 python bench.py > gpurun_out/r1f_bench_c3.json 2> gpurun_out/r1f_bench_c3.err; echo "c3 rc=$?"
timeout 100 python bench.py --workload c2 --no-cpu-baseline > gpurun_out/r1f_bench_c2.json 2> gpurun_out/r1f_bench_c2.err; echo "c2 rc=$?"
timeout 150 python bench.py --workload c4 --steps 5 --no-cpu-baseline > gpurun_out/r1f_bench_c4_1gpu.json 2> gpurun_out/r1f_bench_c4.err; echo "c4 rc=$?"
timeout 150 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r1f_bench_reference.json 2> gpurun_out/r1f_bench_reference.err; echo "ref rc=$?"
python - <<'PY'
import json
for w in ("c3", "c2", "c4_1gpu", "reference"):
    try:
        js = json.load(open("gpurun_out/r1f_bench_%s.json" % w))
        print(w, "%.4g" % js["value"], "%.3f ms" % js["ms_per_step"], "e2e %.4g" % js.get("e2e", {}).get("value", 0),
              {k: v["ms"] for k, v in js.get("roofline", {}).get("passes", {}).items()})
    except Exception as e:
        print(w, "failed", e)
PY
