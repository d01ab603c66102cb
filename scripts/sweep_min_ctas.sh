set -e
cd hc-spmm_b200
for m in 2 3 4; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -DHCSPMM_MIN_CTAS=$m -o lib/libhcspmm.so csrc/capi.cu csrc/preprocess.cu csrc/spmm.cu csrc/gemm.cu
  for lr in 256 1024 4096; do
    echo "== MIN_CTAS=$m long_row=$lr"
    python ../bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --long-row $lr 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'])"
  done
done
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -o lib/libhcspmm.so csrc/capi.cu csrc/preprocess.cu csrc/spmm.cu csrc/gemm.cu
cd ..
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
