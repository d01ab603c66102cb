#!/bin/bash
# round 2, 2 GPUs: row-block pipeline -- tests, products sweep over blocks, bench
mkdir -p gpurun_out
export HCSPMM_TEST_REPORT=gpurun_out/r2_c14_multi_parity_report.txt
rm -f $HCSPMM_TEST_REPORT
timeout 900 python -m pytest tests/test_gpu_peer.py tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r2_c14_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r2_c14_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout 600 $TR scripts/r2/inplace_sweep.py --refs 0 --row-blocks 1 4 --steps 10 2> gpurun_out/r2_c14_sweep.err | grep '^{' > gpurun_out/r2_rowblock_sweep_2.jsonl; echo "sweep rc=$?"; cat gpurun_out/r2_rowblock_sweep_2.jsonl; tail -3 gpurun_out/r2_c14_sweep.err
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2> gpurun_out/r2_c14_bench_2.err | grep '^{' > gpurun_out/r2_c14_bench_2.json; echo "bench2 rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_c14_bench_2.json").read())
    print("bench2", round(d["ms_per_step"],3), d["config"]["phases"], d["parity"], d["launches_per_step"])
    p=d["extra"]["products"]; print("  products", round(p["ms_per_step"],3), p["phases"], p["parity"], p["launches_per_step"])
except Exception as e:
    print("no line", e); print(open("gpurun_out/r2_c14_bench_2.err").read()[-1500:])
PY
