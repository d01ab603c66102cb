# tensor-core path pipeline depth on the proteins-shape (dense windows) graph (run under gpurun)
set -e
cd hc-spmm_b200
SRC="csrc/capi.cu csrc/preprocess.cu csrc/spmm.cu csrc/gemm.cu"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared"
run() { python ../bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   ms %.3f  GFLOP/s %.0f  tc_windows %d' % (d['ms_per_step'], d['value'], d['config']['tc_windows']))"; }
for st in 2 3 4 6; do
  nvcc $FLAGS -DHCSPMM_TC_STAGES=$st -o lib/libhcspmm.so $SRC
  echo "== TC_STAGES=$st proteins dim256 all_tc"; run --shape proteins --classifier all_tc
  echo "== TC_STAGES=$st proteins dim64 all_tc"; run --shape proteins --classifier all_tc --dim 64
done
nvcc $FLAGS -o lib/libhcspmm.so $SRC
echo "== proteins dim256 all_cuda vec8=0"; run --shape proteins --vec8 0
echo "== proteins dim64 all_cuda vec8=0"; run --shape proteins --vec8 0 --dim 64
echo "== proteins dim256 tf32x2 all_tc"; run --shape proteins --classifier all_tc --precision tf32x2
cd ..
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
