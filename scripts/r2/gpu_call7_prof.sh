#!/bin/bash
# round 2, profiling call: new tests, headline bench, launch list, ncu --set full of the top kernels
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "refit or gemm_tma or dense_plan" > gpurun_out/r2_c7_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_c7_tests.log
timeout 600 python bench.py > gpurun_out/r2_c7_bench.json 2> gpurun_out/r2_c7_bench.err; echo "bench rc=$?"; head -c 400 gpurun_out/r2_c7_bench.json; echo
timeout 300 python benchmarks/gcn_epoch.py --epochs 10 --warmup 3 > gpurun_out/r2_c7_gcn_1.json 2> /dev/null; echo "gcn1 rc=$?"; head -c 300 gpurun_out/r2_c7_gcn_1.json; echo
timeout 300 python benchmarks/gcn_epoch.py --shape proteins --model gin --feat 256 --hidden 256 --classes 112 --classifier b200 --dense --epochs 10 --warmup 3 > gpurun_out/r2_c7_gin_1.json 2> /dev/null; echo "gin1 rc=$?"; head -c 300 gpurun_out/r2_c7_gin_1.json; echo
B="python bench.py --steps 3 --warmup 1 --no-extra --no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_bench.csv $B > gpurun_out/r2_c7_ncu_list.log 2>&1; echo "ncu list rc=$?"
for t in reddit_spmm:spmm_balanced_kernel products_spmm:spmm_balanced_kernel products_spmm_degree:spmm_balanced_kernel gemm_products:update_gemm_tma_kernel fused_proteins:spmm_dense_tma_kernel dense_proteins:spmm_dense_ws_kernel; do
  s=${t%%:*}; k=${t##*:}
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/r2_prof_$s python scripts/r2/prof_target.py $s > gpurun_out/r2_c7_ncu_$s.log 2>&1; echo "ncu $s rc=$?"
  # gpurun_out/ comes back only below 64 MiB: keep the text summary + raw metrics of every capture, the .ncu-rep only when small
  python scripts/summarize_ncu.py gpurun_out/r2_prof_$s.ncu-rep gpurun_out/r2_ncu_$s.txt
  ncu -i gpurun_out/r2_prof_$s.ncu-rep --page raw --csv > gpurun_out/r2_ncu_raw_$s.csv 2>/dev/null
  if [ $(stat -c %s gpurun_out/r2_prof_$s.ncu-rep) -gt 6000000 ]; then rm -f gpurun_out/r2_prof_$s.ncu-rep; fi
done
ls -la gpurun_out/ | tail -30; du -sh gpurun_out
