#!/bin/bash
# One gpurun call on N GPUs ($1): the distributed GPU tests, the bench line with every config, the sub-step profile.
#   /usr/local/graft/bin/gpurun --gpus N --timeout 900 -- 'bash tools/gpu_call_r2_multi.sh N'
N=${1:-2}
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_distributed.py -x -q -m gpu > gpurun_out/r2_pytest_g$N.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_pytest_g$N.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --all-configs > gpurun_out/r2_bench_g$N.json 2> gpurun_out/r2_bench_g$N.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_g$N.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["ms_per_step"], d["value"], {k:v["ms"] for k,v in d["roofline"]["passes"].items()}, d["parity"]["ok"], d["e2e"]["ms_per_step"] if d.get("e2e") else None)
print({k:(round(v["ms_per_step"],3), '%.3g'%v["value"]) for k,v in d.get("configs",{}).items()})
PY
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/halo_profile.py > gpurun_out/r2_halo_g$N.txt 2>&1; tail -15 gpurun_out/r2_halo_g$N.txt
