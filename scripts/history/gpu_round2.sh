python -m pytest tests -m gpu -q 2>&1 | tail -3
run() { python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   ms %.3f  GFLOP/s %.0f' % (d['ms_per_step'], d['value']))"; }
echo "== products dim47"; run --shape products --dim 47
echo "== gcn products 1 GPU"; python benchmarks/gcn_epoch.py --epochs 10 --warmup 3 2>&1 | tail -1
echo "== gcn products 1 GPU dense(tcgen05 gemm)"; python - <<'PY'
import sys; sys.path[:0]=['.','hc-spmm_b200']
import subprocess
PY
python scripts/bench_gemm.py 2>&1 | tail -8
python benchmarks/format_sweep.py --quick 2>&1 | tail -8
