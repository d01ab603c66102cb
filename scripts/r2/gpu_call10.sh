#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_module.py tests/test_gpu_loa.py tests/test_gpu_peer.py -x -q -m gpu > gpurun_out/r2_c10_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2_c10_tests.log
timeout 400 python scripts/r2/hint_probe.py > gpurun_out/r2_c10_hint_probe.log 2>&1; echo "hint probe rc=$?"; cat gpurun_out/r2_c10_hint_probe.log | cut -c 1-700
timeout 300 python benchmarks/gcn_epoch.py --epochs 10 --warmup 3 --profile > gpurun_out/r2_c10_gcn_1.json 2> gpurun_out/r2_c10_gcn_1.prof; echo "gcn1 rc=$?"; head -c 250 gpurun_out/r2_c10_gcn_1.json; echo; grep -v Warning gpurun_out/r2_c10_gcn_1.prof | head -40 | cut -c 1-220
timeout 300 python benchmarks/gcn_epoch.py --shape proteins --model gin --feat 256 --hidden 256 --classes 112 --classifier b200 --dense --epochs 10 --warmup 3 --profile > gpurun_out/r2_c10_gin_1.json 2> gpurun_out/r2_c10_gin_1.prof; echo "gin1 rc=$?"; head -c 250 gpurun_out/r2_c10_gin_1.json; echo; grep -v Warning gpurun_out/r2_c10_gin_1.prof | head -32 | cut -c 1-220
timeout 400 python scripts/r2/l2_and_loa.py --loa 2>&1 | grep loa
B="python bench.py --steps 3 --warmup 1 --no-extra --no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"spmm|merge_path|preprocess|fixup|tag_columns|rank_class|count_columns" -c 40 --csv --log-file gpurun_out/r2_launches_bench.csv $B > gpurun_out/r2_c10_ncu_list.log 2>&1; echo "ncu list rc=$?"
