python -m pytest tests -m gpu -q -x 2>&1 | tail -8
run() { python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   ms %.3f  GFLOP/s %.0f' % (d['ms_per_step'], d['value']))"; }
echo "== envelope dim32"; run --shape envelope --dim 32
echo "== products dim32"; run --shape products --dim 32
echo "== products dim128"; run --shape products
echo "== products dim47"; run --shape products --dim 47
echo "== gcn products 1 GPU"; python benchmarks/gcn_epoch.py --epochs 10 --warmup 3 2>&1 | tail -1
