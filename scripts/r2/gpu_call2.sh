#!/bin/bash
# round 2, call 2: full GPU suite, smoke, bench N=1 (new line), launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_c2_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2_c2_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_c2_smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/r2_c2_smoke.log
timeout 600 python bench.py > gpurun_out/r2_c2_bench.json 2> gpurun_out/r2_c2_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_c2_bench.err; head -c 3000 gpurun_out/r2_c2_bench.json
