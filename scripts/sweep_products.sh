set -e
cd hc-spmm_b200
SRC="csrc/capi.cu csrc/preprocess.cu csrc/spmm.cu csrc/gemm.cu csrc/loa.cu csrc/umma_gemm.cu csrc/dense.cu"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared"
run() { python ../bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   ms %.3f  GFLOP/s %.0f' % (d['ms_per_step'], d['value']))"; }
for m in 2 3 4; do for inf in 128 256; do
  nvcc $FLAGS -DHCSPMM_MIN_CTAS=$m -DHCSPMM_INFLIGHT_BYTES=$inf -o lib/libhcspmm.so $SRC
  for w in 1 2 4; do
    echo "== MIN_CTAS=$m INFLIGHT=$inf wpc=$w products dim128"; run --shape products --tune wpc=$w
  done
  echo "== MIN_CTAS=$m INFLIGHT=$inf products dim64"; run --shape products --dim 64
  echo "== MIN_CTAS=$m INFLIGHT=$inf reddit dim256"; run
done; done
nvcc $FLAGS -o lib/libhcspmm.so $SRC
