# one GPU visit: full parity suite, smoke, headline bench (+reference arm), ncu launch list + full capture of the balanced kernel
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py > gpurun_out/bench_reddit_r1b.json 2> gpurun_out/bench_reddit_r1b.err; cut -c1-2500 gpurun_out/bench_reddit_r1b.json
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_reference_r1b.json; cut -c1-600 gpurun_out/bench_reference_r1b.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_r1b.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:spmm_balanced_kernel -s 3 -c 1 -o gpurun_out/prof_spmm_balanced $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_r1b.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[-12:]: print(f"{float(r[-1])/1e3:9.1f} us  {r[4][:80]}  grid {r[8]}")
PY
