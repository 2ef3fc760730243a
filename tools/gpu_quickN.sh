# N GPUs ($1): distributed tests, then the bench line without and with exchange B overlapped
N=${1:-2}
timeout 500 python -m pytest tests/test_gpu_distributed.py -x -q -m gpu > gpurun_out/qN_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/qN_pytest.log
for ov in 0 1; do
SPH_OVERLAP_B=$ov timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --no-cpu-baseline --no-e2e > gpurun_out/qN_bench_$ov.json 2> gpurun_out/qN_bench_$ov.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/qN_bench_$ov.json").read().strip().splitlines()[-1])
print("overlap_b=$ov", d["n_gpus"], d["ms_per_step"], {k:v["ms"] for k,v in d["roofline"]["passes"].items()}, d["parity"]["ok"])
PY
done
