#!/bin/bash
# round 2, call 1: TMA GEMM tests + probe, L2 roof, LOA timing, full GPU suite
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "gemm" > gpurun_out/r2_c1_gemm_tests.log 2>&1; echo "gemm tests rc=$?"
tail -5 gpurun_out/r2_c1_gemm_tests.log
timeout 300 python scripts/r2/gemm_probe.py > gpurun_out/r2_c1_gemm_probe.log 2>&1; echo "probe rc=$?"
cat gpurun_out/r2_c1_gemm_probe.log | tail -30
timeout 400 python scripts/r2/l2_and_loa.py --loa > gpurun_out/r2_c1_l2_loa.log 2>&1; echo "l2/loa rc=$?"
cat gpurun_out/r2_c1_l2_loa.log | tail -20
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r2_c1_all_tests.log 2>&1; echo "all tests rc=$?"
tail -3 gpurun_out/r2_c1_all_tests.log
