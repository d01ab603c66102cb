#!/bin/bash
# round 2, 2 GPUs: last check of the defaults (multi-GPU parity tests, bench)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_peer.py -q -m gpu > gpurun_out/r2_c35_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r2_c35_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 2> gpurun_out/r2_c35_bench_2.err | grep '^{' > gpurun_out/r2_c35_bench_2.json; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_c35_bench_2.json").read())
print("bench2", round(d["ms_per_step"],4), d["parity"], d["launches_per_step"], d["e2e"]["ms_per_step"])
p=d["extra"]["products"]; print("  products", round(p["ms_per_step"],4), p["phases"], p["parity"]["rel_fro"], p["launches_per_step"])
PY
