#!/bin/bash
# round 2, 2 GPUs: multi-process parity test, bench at N=2 (fp32 + bf16 operand), GCN epoch at N=2
mkdir -p gpurun_out
export HCSPMM_TEST_REPORT=gpurun_out/r2_multi_parity_report.txt
rm -f $HCSPMM_TEST_REPORT
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r2_c6_multi_tests.log 2>&1; echo "multi tests rc=$?"; tail -5 gpurun_out/r2_c6_multi_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_c6_bench_2.json 2> gpurun_out/r2_c6_bench_2.err; echo "bench2 rc=$?"; tail -2 gpurun_out/r2_c6_bench_2.err; head -c 1500 gpurun_out/r2_c6_bench_2.json; echo
timeout 600 $TR bench.py --gpus 2 --steps 10 --warmup 3 --precision bf16 > gpurun_out/r2_c6_bench_2_bf16.json 2> gpurun_out/r2_c6_bench_2_bf16.err; echo "bench2 bf16 rc=$?"; tail -2 gpurun_out/r2_c6_bench_2_bf16.err; head -c 600 gpurun_out/r2_c6_bench_2_bf16.json; echo
timeout 600 $TR benchmarks/gcn_epoch.py --epochs 10 --warmup 3 > gpurun_out/r2_c6_gcn_2.json 2> gpurun_out/r2_c6_gcn_2.err; echo "gcn2 rc=$?"; tail -2 gpurun_out/r2_c6_gcn_2.err; cat gpurun_out/r2_c6_gcn_2.json
timeout 600 python benchmarks/gcn_epoch.py --epochs 10 --warmup 3 > gpurun_out/r2_c6_gcn_1.json 2> gpurun_out/r2_c6_gcn_1.err; echo "gcn1 rc=$?"; cat gpurun_out/r2_c6_gcn_1.json
