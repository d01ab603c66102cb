set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -6
run() { python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   ms %.3f  GFLOP/s %.0f' % (d['ms_per_step'], d['value']))"; }
echo "== envelope dim32"; run --shape envelope --dim 32
echo "== envelope dim32 vec8=0"; run --shape envelope --dim 32 --vec8 0
echo "== envelope dim64"; run --shape envelope --dim 64
echo "== envelope dim128"; run --shape envelope --dim 128
echo "== products dim128"; run --shape products
echo "== products dim32"; run --shape products --dim 32
echo "== products dim64"; run --shape products --dim 64
echo "== reddit dim256"; run
echo "== reddit dim32"; run --dim 32
python bench.py --impl reference --ref-kernel --shape envelope --steps 50 --warmup 5 2>/dev/null | tail -1 | cut -c1-200
