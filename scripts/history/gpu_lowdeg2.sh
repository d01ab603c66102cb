run() { python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   ms %.3f  GFLOP/s %.0f' % (d['ms_per_step'], d['value']))"; }
python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2
for sr in 0 2 8 32; do for w in 0 4 8; do echo "== envelope dim32 short_row=$sr wpc=$w"; run --shape envelope --dim 32 --tune short_row=$sr --tune wpc=$w; done; done
for sr in 0 4 8 16; do echo "== products dim128 short_row=$sr"; run --shape products --tune short_row=$sr; done
for sr in 0 8; do echo "== products dim32 short_row=$sr"; run --shape products --dim 32 --tune short_row=$sr; done
for sr in 0 8; do echo "== products dim64 short_row=$sr"; run --shape products --dim 64 --tune short_row=$sr; done
echo "== envelope dim64"; run --shape envelope --dim 64
echo "== envelope dim128"; run --shape envelope --dim 128
