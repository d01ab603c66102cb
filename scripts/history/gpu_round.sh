# one GPU visit: parity tests, smoke, headline bench, reference arms, ncu launch list + full capture
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -4
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py > gpurun_out/bench_reddit.json 2> gpurun_out/bench_reddit.err; cat gpurun_out/bench_reddit.json
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_reference.json; cat gpurun_out/bench_reference.json
python bench.py --shape envelope --dim 32 --steps 50 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 > gpurun_out/bench_envelope_ours.json; cat gpurun_out/bench_envelope_ours.json
python bench.py --impl reference --ref-kernel --shape envelope --steps 50 --warmup 5 2>/dev/null | tail -1 > gpurun_out/bench_envelope_refkernel.json; cat gpurun_out/bench_envelope_refkernel.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:spmm_hybrid -s 3 -c 1 -o gpurun_out/prof_spmm $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
