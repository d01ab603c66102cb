#!/bin/bash
# round 2, 8 GPUs, ONE build: bench (Reddit headline + products sub-record), in-place sweep on the products shape,
# GCN / GIN epochs, real-peer parity at world 8, feature-slab overlap, peer-read roof, config-5 sweep at 500 M entries
mkdir -p gpurun_out
run() { # n, name, timeout, script args...
  n=$1; name=$2; to=$3; shift; shift; shift
  timeout $to python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 "$@" 2> gpurun_out/r2_c13_$name.err | grep '^{' > gpurun_out/r2_c13_$name.json
  echo "$name rc=$? $(head -c 300 gpurun_out/r2_c13_$name.json)"
}
run 8 bench_8 240 bench.py --gpus 8 --steps 20 --warmup 5
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_c13_bench_8.json").read())
    print("bench8", round(d["ms_per_step"],4), d["value"], d["config"]["phases"], d["parity"], (d.get("e2e") or {}).get("ms_per_step"))
    p=d["extra"]["products"]; print("  products", round(p["ms_per_step"],4), p["phases"], p["parity"])
except Exception as e:
    print("no line", e); print(open("gpurun_out/r2_c13_bench_8.err").read()[-2000:])
PY
run 8 inplace_8 300 scripts/r2/inplace_sweep.py --refs 0 1 2 3 4 --steps 20
cat gpurun_out/r2_c13_inplace_8.json; tail -2 gpurun_out/r2_c13_inplace_8.err
run 8 gcn_8 200 benchmarks/gcn_epoch.py --epochs 10 --warmup 3
run 8 gin_8 200 benchmarks/gcn_epoch.py --shape proteins --model gin --feat 256 --hidden 256 --classes 112 --classifier b200 --dense --epochs 10 --warmup 3
export HCSPMM_TEST_REPORT=gpurun_out/r2_multi_parity_report_8.txt
rm -f $HCSPMM_TEST_REPORT
timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -m gpu -k "8" > gpurun_out/r2_c13_multi_tests.log 2>&1; echo "multi tests rc=$?"; tail -3 gpurun_out/r2_c13_multi_tests.log
run 8 bench_8_slabs2 150 bench.py --gpus 8 --steps 20 --warmup 5 --exchange-slabs 2 --no-extra --no-e2e --no-cpu-baseline
run 8 bench_8_gather 150 bench.py --gpus 8 --steps 10 --warmup 3 --exchange gather --no-extra --no-e2e --no-cpu-baseline
run 8 bw_8 100 scripts/r2/peer_bw_probe.py
cat gpurun_out/r2_c13_bw_8.json
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 benchmarks/format_sweep.py --nnz 500000000 --deg 128 --dims 32 256 --bands 256 --skip-all-tc > gpurun_out/r2_format_sweep_8gpu.jsonl 2> gpurun_out/r2_c13_sweep.err; echo "sweep rc=$?"; tail -2 gpurun_out/r2_c13_sweep.err; cut -c 1-400 gpurun_out/r2_format_sweep_8gpu.jsonl
