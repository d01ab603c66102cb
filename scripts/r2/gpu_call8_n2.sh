#!/bin/bash
# round 2, 2 GPUs (cheap check before the 8-GPU call): push schedule parity + timing, LOA tests, quick sweep under torchrun
mkdir -p gpurun_out
export HCSPMM_TEST_REPORT=gpurun_out/r2_multi_parity_report.txt
rm -f $HCSPMM_TEST_REPORT
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_peer.py tests/test_gpu_loa.py -x -q -m gpu > gpurun_out/r2_c8_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2_c8_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
for ex in peer push; do
  timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 --exchange $ex --no-cpu-baseline --no-e2e 2> gpurun_out/r2_c8_bench_2_$ex.err | grep '^{' > gpurun_out/r2_c8_bench_2_$ex.json; echo "bench2 $ex rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2_c8_bench_2_$ex.json").read())
print("$ex", round(d["ms_per_step"],3), d["config"]["phases"], d["parity"]["rel_fro"])
p=d["extra"]["products"]; print("  products", round(p["ms_per_step"],3), p["phases"], p["parity"]["rel_fro"])
PY
done
timeout 600 $TR benchmarks/format_sweep.py --nnz 10000000 --deg 128 --dims 32 256 --bands 256 --skip-all-tc > gpurun_out/r2_c8_sweep_2.jsonl 2> gpurun_out/r2_c8_sweep_2.err; echo "sweep2 rc=$?"; tail -2 gpurun_out/r2_c8_sweep_2.err; cut -c 1-400 gpurun_out/r2_c8_sweep_2.jsonl
timeout 300 python scripts/r2/l2_and_loa.py --loa 2>&1 | grep loa
