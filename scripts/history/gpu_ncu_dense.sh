CMD="python bench.py --shape proteins --classifier all_tc --dense --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_dense.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:spmm_dense -s 2 -c 1 -o gpurun_out/prof_dense $CMD > gpurun_out/ncu_dense.log 2>&1
tail -2 gpurun_out/ncu_dense.log
$CMD > gpurun_out/plain_dense2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_dense.csv $CMD > /dev/null 2>&1
grep -E "spmm_dense|tf32_round|spmm_hybrid|preprocess|dense_" gpurun_out/launches_dense.csv | awk -F'","' '{print $5, $NF}' | tail -14
