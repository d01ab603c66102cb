N=${N:-2}
python -m pytest tests/test_gpu_peer.py -m gpu -q 2>&1 | tail -2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR scripts/gpu_peer_check.py 2>&1 | grep -vE "Warning|warn|^\s*$|\*\*\*" | tail -8
show() { tee gpurun_out/last_multi.log | grep -E '^\{' | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('  ms/step %.3f  GFLOP/s %.0f  e2e %s  phases %s  exchange: %s' % (d['ms_per_step'], d['value'], d['e2e'] and round(d['e2e']['value']), d['config'].get('phases'), d['config']['exchange'][:60]))" || tail -20 gpurun_out/last_multi.log; }
echo "== bench products $N GPUs exchange=peer"; timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --shape products --exchange peer --no-e2e 2>&1 | show
echo "== bench reddit $N GPUs auto"; timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 2>&1 | show
echo "== gcn $N GPUs peer"; timeout 300 $TR benchmarks/gcn_epoch.py --epochs 8 --warmup 3 --schedule peer > gpurun_out/last_gcn.log 2>&1; grep -E '^\{' gpurun_out/last_gcn.log | tail -1 | cut -c1-1300; grep -E '^\{' -q gpurun_out/last_gcn.log || grep -vE "Warning|warn" gpurun_out/last_gcn.log | tail -25
