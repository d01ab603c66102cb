#!/bin/bash
# round 2, 2 GPUs: segment mode (rows read in place from the peers' operands) -- single-device kernel test, real-peer
# parity, bench with / without it (FP32 and BF16 operand), peer-read roof
mkdir -p gpurun_out
export HCSPMM_TEST_REPORT=gpurun_out/r2_c11_multi_parity_report.txt
rm -f $HCSPMM_TEST_REPORT
timeout 900 python -m pytest tests/test_gpu_peer.py tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r2_c11_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2_c11_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout 300 $TR scripts/r2/peer_bw_probe.py 2> gpurun_out/r2_c11_bw.err | grep '^{' > gpurun_out/r2_peer_bw_probe_2.json; echo "bw rc=$?"; cat gpurun_out/r2_peer_bw_probe_2.json
for v in "auto:" "pull:--direct-refs 0" "t1:--direct-refs 1" "t4:--direct-refs 4" "bf16:--precision bf16" "bf16pull:--precision bf16 --direct-refs 0"; do
  name=${v%%:*}; flags=${v#*:}
  timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e $flags 2> gpurun_out/r2_c11_bench_2_$name.err | grep '^{' > gpurun_out/r2_c11_bench_2_$name.json; echo "bench2 $name rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_c11_bench_2_$name.json").read())
    print("$name", round(d["ms_per_step"],3), d["config"]["phases"], d["parity"])
    p=d["extra"]["products"]; print("  products", round(p["ms_per_step"],3), p["phases"], p["parity"])
except Exception as e:
    print("$name: no line", e); print(open("gpurun_out/r2_c11_bench_2_$name.err").read()[-1500:])
PY
done
