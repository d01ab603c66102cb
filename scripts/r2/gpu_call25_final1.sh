#!/bin/bash
# round 2, final 1-GPU call: full GPU suite, headline bench, launch list, ncu --set full of the two balanced-kernel shapes
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2_c25_tests_all.log 2>&1; echo "all tests rc=$?"; tail -3 gpurun_out/r2_c25_tests_all.log | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_c25_bench_1.json 2> gpurun_out/r2_c25_bench_1.err; echo "bench rc=$?"; head -c 300 gpurun_out/r2_c25_bench_1.json; echo
timeout 300 python benchmarks/gcn_epoch.py --epochs 10 --warmup 3 > gpurun_out/r2_c25_gcn_1.json 2> /dev/null; echo "gcn1 rc=$?"; head -c 120 gpurun_out/r2_c25_gcn_1.json; echo
B="python bench.py --steps 3 --warmup 1 --no-extra --no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"spmm|merge_path|preprocess|fixup|rowsort" -c 40 --csv --log-file gpurun_out/r2_launches_bench.csv $B > gpurun_out/r2_c25_ncu_list.log 2>&1; echo "ncu list rc=$?"
for t in reddit_spmm:spmm_balanced_kernel products_spmm:spmm_balanced_kernel; do
  s=${t%%:*}; k=${t##*:}
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o gpurun_out/r2_prof_$s python scripts/r2/prof_target.py $s > gpurun_out/r2_c25_ncu_$s.log 2>&1; echo "ncu $s rc=$?"
  python scripts/summarize_ncu.py gpurun_out/r2_prof_$s.ncu-rep gpurun_out/r2_ncu_$s.txt
  if [ $(stat -c %s gpurun_out/r2_prof_$s.ncu-rep) -gt 6000000 ]; then rm -f gpurun_out/r2_prof_$s.ncu-rep; fi
done
cat gpurun_out/r2_ncu_products_spmm.txt | head -30
