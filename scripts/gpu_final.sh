# one GPU visit at the end of a round: full parity suite, smoke, headline bench + reference arm, products ncu capture
python -m pytest tests -m gpu -q 2>&1 | tail -4
python __graft_entry__.py smoke 2>&1 | tail -2
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; cut -c1-400 gpurun_out/bench_final.json; python -c "
import json; d=json.load(open('gpurun_out/bench_final.json')); print('value', d['value'], 'ms', d['ms_per_step'], 'roofline', d['roofline']['frac'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], 'clocks', d['clocks'])"
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
CMD="python bench.py --shape products --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_prod.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:spmm_balanced_kernel -s 3 -c 1 -o gpurun_out/prof_products_balanced $CMD > gpurun_out/ncu_prod.log 2>&1
tail -1 gpurun_out/ncu_prod.log; tail -1 gpurun_out/plain_prod.log | cut -c1-300
