#!/bin/bash
mkdir -p gpurun_out
timeout 900 python benchmarks/selector_fit.py > gpurun_out/r2_c5_selector_fit.log 2>&1; echo "selector fit rc=$?"; tail -40 gpurun_out/r2_c5_selector_fit.log
timeout 1200 python scripts/r2/locality.py --products-loa 700 > gpurun_out/r2_c5_locality.log 2>&1; echo "locality rc=$?"; cat gpurun_out/r2_c5_locality.log | tail -12
timeout 200 python scripts/r2/dense_probe.py > gpurun_out/r2_c5_dense_probe.log 2>&1; echo "probe rc=$?"; tail -4 gpurun_out/r2_c5_dense_probe.log | cut -c 1-1200
