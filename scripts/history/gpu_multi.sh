# 2-GPU validation: bench (strong scaling SpMM with all-gather) and GCN epoch, both schedules
run() { python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   ms %.3f  GFLOP/s %.0f' % (d['ms_per_step'], d['value']))"; }
for sr in 0 1; do for d in 32 64 128; do echo "== products dim$d short_row=$sr"; run --shape products --dim $d --tune short_row=$sr; done; done
echo "== envelope dim32"; run --shape envelope --dim 32
N=${NGPU:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== bench 1 GPU"; python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-400
echo "== bench $N GPUs"; $TR bench.py --gpus $N --steps 20 --warmup 5 2>&1 | grep -E '^\{' | tail -1 | cut -c1-1200
echo "== bench reference arm under torchrun"; $TR bench.py --gpus $N --impl reference --steps 2 --warmup 1 2>&1 | grep -E '^\{' | tail -1 | cut -c1-300
echo "== gcn products 1 GPU"; python benchmarks/gcn_epoch.py --epochs 10 --warmup 3 2>&1 | tail -1
echo "== gcn products $N GPUs gather"; $TR benchmarks/gcn_epoch.py --epochs 10 --warmup 3 2>&1 | grep -E '^\{' | tail -1
echo "== gcn products $N GPUs slabs"; $TR benchmarks/gcn_epoch.py --epochs 10 --warmup 3 --schedule slabs 2>&1 | grep -E '^\{|Error|error' | tail -3
