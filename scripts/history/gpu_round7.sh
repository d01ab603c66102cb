python -m pytest tests -m gpu -q 2>&1 | tail -3
run() { timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e "$@" 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   ms %.3f  GFLOP/s %.0f' % (d['ms_per_step'], d['value']))"; }
echo "== reddit dim256 fp32"; run
echo "== reddit dim256 bf16"; run --precision bf16
echo "== reddit dim128 bf16"; run --precision bf16 --dim 128
echo "== reddit dim512 bf16"; run --precision bf16 --dim 512
echo "== products dim128 bf16"; run --precision bf16 --shape products
echo "== proteins dim256 bf16"; run --precision bf16 --shape proteins
