python -m pytest tests -m gpu -q 2>&1 | tail -3
run() { python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   ms %.3f  GFLOP/s %.0f' % (d['ms_per_step'], d['value']))"; }
echo "== products dim47"; run --shape products --dim 47
echo "== products dim48"; run --shape products --dim 48
echo "== products dim96"; run --shape products --dim 96
echo "== products dim100"; run --shape products --dim 100
echo "== reddit dim256"; run
echo "== reddit dim200"; run --dim 200
CMD="python bench.py --shape proteins --classifier all_tc --dense --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_dense.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:spmm_dense -s 2 -c 1 -o gpurun_out/prof_dense2 $CMD > gpurun_out/ncu_dense.log 2>&1
tail -1 gpurun_out/ncu_dense.log
CMD="python bench.py --shape products --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_prod.log 2>&1 && ncu --set full --clock-control none -k regex:spmm_hybrid -s 3 -c 1 -o gpurun_out/prof_products $CMD > gpurun_out/ncu_prod.log 2>&1
tail -1 gpurun_out/ncu_prod.log
