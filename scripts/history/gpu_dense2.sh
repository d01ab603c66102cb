python -m pytest tests -m gpu -q -x 2>&1 | tail -3
run() { python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   ms %.3f  GFLOP/s %.0f  tc_windows %d dense_groups %d prep_ms %.2f' % (d['ms_per_step'], d['value'], d['config']['tc_windows'], d['config']['dense_groups_tcgen05'], d['config']['preprocess_ms']))"; }
for d in 256 128 64; do
echo "== proteins dim$d all CUDA-core"; run --shape proteins --dim $d
echo "== proteins dim$d all_tc mma.sync"; run --shape proteins --dim $d --classifier all_tc
echo "== proteins dim$d all_tc + tcgen05 dense"; run --shape proteins --dim $d --classifier all_tc --dense
echo "== proteins dim$d b200 + tcgen05 dense"; run --shape proteins --dim $d --classifier b200 --dense
done
echo "== reddit dim256 b200 + dense (expect no dense groups)"; run --classifier b200 --dense
