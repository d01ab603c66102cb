python -m pytest tests -m gpu -q 2>&1 | tail -3
run() { python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   ms %.3f  GFLOP/s %.0f' % (d['ms_per_step'], d['value']))"; }
echo "== products dim128"; run --shape products
echo "== products dim128 occupancy3=0"; run --shape products --tune occupancy3=0
echo "== products dim64"; run --shape products --dim 64
echo "== products dim32"; run --shape products --dim 32
echo "== products dim47"; run --shape products --dim 47
echo "== envelope dim32"; run --shape envelope --dim 32
echo "== reddit dim256"; run
echo "== proteins all_tc dense v1"; run --shape proteins --classifier all_tc --dense
echo "== proteins all_tc dense ws"; timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --shape proteins --classifier all_tc --dense --tune dense_ws=1 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   ms %.3f  GFLOP/s %.0f' % (d['ms_per_step'], d['value']))"
echo "== proteins dim128 dense ws"; timeout 120 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --shape proteins --dim 128 --classifier all_tc --dense --tune dense_ws=1 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   ms %.3f  GFLOP/s %.0f' % (d['ms_per_step'], d['value']))"
echo "== gcn products 1 GPU"; python benchmarks/gcn_epoch.py --epochs 10 --warmup 3 2>&1 | tail -1 | cut -c1-120
