# strong-scaling runs on one box: bench.py (SpMM + all-gather) and the products-shape GCN epoch
NMAX=${NMAX:-8}
echo "== bench 1 GPU"; python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/scale_bench_1.json; python -c "import json; d=json.load(open('gpurun_out/scale_bench_1.json')); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'])"
echo "== gcn 1 GPU"; python benchmarks/gcn_epoch.py --epochs 8 --warmup 3 2>/dev/null | tail -1 > gpurun_out/scale_gcn_1.json; cut -c1-90 gpurun_out/scale_gcn_1.json
for N in 2 4 8; do
  [ $N -le $NMAX ] || continue
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N"
  echo "== bench $N GPUs"; $TR bench.py --gpus $N --steps 10 --warmup 3 2>/dev/null | grep -E '^\{' | tail -1 > gpurun_out/scale_bench_$N.json; python -c "import json; d=json.load(open('gpurun_out/scale_bench_$N.json')); print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'])"
  echo "== gcn $N GPUs gather"; $TR benchmarks/gcn_epoch.py --epochs 8 --warmup 3 2>/dev/null | grep -E '^\{' | tail -1 > gpurun_out/scale_gcn_$N.json; cut -c1-90 gpurun_out/scale_gcn_$N.json
  echo "== gcn $N GPUs slabs"; $TR benchmarks/gcn_epoch.py --epochs 8 --warmup 3 --schedule slabs 2>/dev/null | grep -E '^\{' | tail -1 > gpurun_out/scale_gcn_slabs_$N.json; cut -c1-90 gpurun_out/scale_gcn_slabs_$N.json
done
