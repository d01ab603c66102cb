# sweep of the CUDA-core gather's compile-time knobs on the headline shape (run under gpurun)
set -e
cd hc-spmm_b200
SRC="csrc/capi.cu csrc/preprocess.cu csrc/spmm.cu csrc/gemm.cu"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared"
run() { python ../bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   ms %.3f  GFLOP/s %.0f' % (d['ms_per_step'], d['value']))"; }
for m in 1 2; do for inf in 128 256; do
  nvcc $FLAGS -DHCSPMM_MIN_CTAS=$m -DHCSPMM_INFLIGHT_BYTES=$inf -o lib/libhcspmm.so $SRC
  for v in 0 1; do
    echo "== MIN_CTAS=$m INFLIGHT=$inf vec8=$v reddit dim256"; run --vec8 $v
  done
done; done
nvcc $FLAGS -o lib/libhcspmm.so $SRC
for d in 32 64 128 512; do for v in 0 1; do echo "== default build vec8=$v reddit dim$d"; run --vec8 $v --dim $d; done; done
for v in 0 1; do echo "== default build vec8=$v products dim128"; run --vec8 $v --shape products; done
for v in 0 1; do echo "== default build vec8=$v proteins dim256 (all CUDA-core)"; run --vec8 $v --shape proteins; done
echo "== proteins dim256 all_tc"; run --shape proteins --classifier all_tc
echo "== proteins dim256 b200 selector"; run --shape proteins --classifier b200
cd ..
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
