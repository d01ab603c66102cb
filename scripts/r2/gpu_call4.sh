#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "dense or fused" > gpurun_out/r2_c4_dense_tests.log 2>&1; echo "dense tests rc=$?"
tail -8 gpurun_out/r2_c4_dense_tests.log
timeout 300 python scripts/r2/dense_probe.py > gpurun_out/r2_c4_dense_probe.log 2>&1; echo "probe rc=$?"; tail -8 gpurun_out/r2_c4_dense_probe.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_c4_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2_c4_tests.log
