for d in 32 256; do
CMD="python bench.py --shape proteins --dim $d --classifier all_tc --dense --tune dense_ws=1 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_d.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_dense_ws_$d.csv $CMD > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_dense_ws_$d.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[-9:]: print(f"dim$d {float(r[-1])/1e3:9.1f} us  {r[4][:70]}  grid {r[8]}")
PY
done
CMD="python bench.py --shape proteins --classifier all_tc --dense --tune dense_ws=1 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_d.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:spmm_dense_ws -s 2 -c 1 -o gpurun_out/prof_dense_ws $CMD > gpurun_out/ncu_dws.log 2>&1
tail -1 gpurun_out/ncu_dws.log
