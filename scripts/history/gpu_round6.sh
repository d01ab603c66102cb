timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "dense" 2>&1 | tail -2
run() { timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e "$@" 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   ms %.3f  GFLOP/s %.0f' % (d['ms_per_step'], d['value']))"; }
for d in 256 128 64; do
echo "== proteins dim$d dense ws (512 producers)"; run --shape proteins --dim $d --classifier all_tc --dense --tune dense_ws=1
done
