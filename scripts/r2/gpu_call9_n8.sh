#!/bin/bash
# round 2, 8 GPUs: multi-process parity (world 2/4/8), bench at 8 and 4 (Reddit headline + products sub-record) with the
# pull / push / gather schedules and the BF16 operand, GCN / GIN epochs at 8 and 4, config-5 sweep up to 500 M entries
mkdir -p gpurun_out
EX=${EX:-peer}
export HCSPMM_TEST_REPORT=gpurun_out/r2_multi_parity_report_8.txt
rm -f $HCSPMM_TEST_REPORT
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r2_c9_multi_tests.log 2>&1; echo "multi tests rc=$?"; tail -3 gpurun_out/r2_c9_multi_tests.log
run() { # n, name, script args...
  n=$1; name=$2; shift; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 "$@" 2> gpurun_out/r2_c9_$name.err | grep '^{' > gpurun_out/r2_c9_$name.json
  echo "$name rc=$? $(head -c 230 gpurun_out/r2_c9_$name.json)"
}
run 8 bench_8 bench.py --gpus 8 --steps 20 --warmup 5
run 8 bench_8_push bench.py --gpus 8 --steps 20 --warmup 5 --exchange push --no-cpu-baseline --no-e2e
run 8 bench_8_gather bench.py --gpus 8 --steps 10 --warmup 3 --exchange gather --no-cpu-baseline --no-e2e
run 8 bench_8_bf16 bench.py --gpus 8 --steps 20 --warmup 5 --precision bf16 --no-cpu-baseline --no-e2e
run 4 bench_4 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e
run 8 gcn_8 benchmarks/gcn_epoch.py --epochs 20 --warmup 5 --schedule $EX
run 4 gcn_4 benchmarks/gcn_epoch.py --epochs 20 --warmup 5 --schedule $EX
run 8 gin_8 benchmarks/gcn_epoch.py --shape proteins --model gin --feat 256 --hidden 256 --classes 112 --classifier b200 --dense --epochs 20 --warmup 5 --schedule $EX
run 4 gin_4 benchmarks/gcn_epoch.py --shape proteins --model gin --feat 256 --hidden 256 --classes 112 --classifier b200 --dense --epochs 20 --warmup 5 --schedule $EX
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 benchmarks/format_sweep.py --nnz 100000000 500000000 --deg 128 --dims 32 256 --bands 256 --skip-all-tc > gpurun_out/r2_format_sweep_8gpu.jsonl 2> gpurun_out/r2_c9_sweep.err; echo "sweep rc=$?"; tail -2 gpurun_out/r2_c9_sweep.err; cut -c 1-300 gpurun_out/r2_format_sweep_8gpu.jsonl
