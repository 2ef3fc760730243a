#!/bin/bash
# One short gpurun call (one GPU, ~45 s of box time): the neighbour-kernel and parity tests, then a 10-step bench line
# with every config -- the edit/measure loop of round 2.
#   /usr/local/graft/bin/gpurun --timeout 600 -- "bash tools/gpu_quick.sh"
timeout 280 python -m pytest tests/test_gpu_tiles.py tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/q_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/q_pytest.log
timeout 200 python bench.py --steps 10 --no-e2e --no-cpu-baseline --all-configs > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/q_bench.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], {k:v["ms"] for k,v in d["roofline"]["passes"].items()}, d["parity"]["ok"], {k:round(v["ms_per_step"],3) for k,v in d["configs"].items()})
PY
