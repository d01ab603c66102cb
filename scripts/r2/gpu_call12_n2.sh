#!/bin/bash
# round 2, 2 GPUs: segment mode + X in the operand's own rows -- tests, T sweep on the products shape, bench, GCN / GIN epochs
mkdir -p gpurun_out
export HCSPMM_TEST_REPORT=gpurun_out/r2_c12_multi_parity_report.txt
rm -f $HCSPMM_TEST_REPORT
timeout 900 python -m pytest tests/test_gpu_peer.py tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r2_c12_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2_c12_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout 600 $TR scripts/r2/inplace_sweep.py --refs 0 2 --steps 10 2> gpurun_out/r2_c12_sweep.err | grep '^{' > gpurun_out/r2_inplace_sweep_2.jsonl; echo "sweep rc=$?"; cat gpurun_out/r2_inplace_sweep_2.jsonl; tail -3 gpurun_out/r2_c12_sweep.err
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/r2_c12_bench_2.err | grep '^{' > gpurun_out/r2_c12_bench_2.json; echo "bench2 rc=$?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_c12_bench_2.json").read())
    print("bench2", round(d["ms_per_step"],3), d["config"]["phases"], d["parity"], d["e2e"])
    p=d["extra"]["products"]; print("  products", round(p["ms_per_step"],3), p["phases"], p["parity"])
except Exception as e:
    print("no line", e); print(open("gpurun_out/r2_c12_bench_2.err").read()[-1500:])
PY
timeout 600 $TR benchmarks/gcn_epoch.py --epochs 10 --warmup 3 2> gpurun_out/r2_c12_gcn_2.err | grep '^{' > gpurun_out/r2_c12_gcn_2.json; echo "gcn2 rc=$?"; cut -c 1-1500 gpurun_out/r2_c12_gcn_2.json; tail -3 gpurun_out/r2_c12_gcn_2.err
timeout 600 $TR benchmarks/gcn_epoch.py --shape proteins --model gin --feat 256 --hidden 256 --classes 112 --classifier b200 --dense --epochs 10 --warmup 3 2> gpurun_out/r2_c12_gin_2.err | grep '^{' > gpurun_out/r2_c12_gin_2.json; echo "gin2 rc=$?"; cut -c 1-1500 gpurun_out/r2_c12_gin_2.json; tail -3 gpurun_out/r2_c12_gin_2.err
