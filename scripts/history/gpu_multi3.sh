S="python benchmarks/sweep_kernel.py"
$S --shape proteins --set balance=0 --set balance=1 2>&1 | grep '^{' | cut -c1-300
$S --shape reddit --set balance=0 --set balance=1 2>&1 | grep '^{' | cut -c1-300
N=${N:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
show() { grep -E '^\{' | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('  ms/step %.3f  GFLOP/s %.0f  e2e %s  phases %s  exchange: %s' % (d['ms_per_step'], d['value'], d['e2e'] and round(d['e2e']['value']), d['config'].get('phases'), d['config']['exchange']))"; }
for EX in gather halo; do
  echo "== bench products $N GPUs exchange=$EX"; $TR bench.py --gpus $N --steps 10 --warmup 3 --shape products --exchange $EX --no-e2e 2>&1 | show
done
echo "== bench products $N GPUs exchange=halo slabs=2"; $TR bench.py --gpus $N --steps 10 --warmup 3 --shape products --exchange halo --exchange-slabs 2 --no-e2e 2>&1 | show
echo "== bench reddit $N GPUs (default)"; $TR bench.py --gpus $N --steps 10 --warmup 3 2>&1 | show
echo "== bench reddit $N GPUs slabs=2"; $TR bench.py --gpus $N --steps 10 --warmup 3 --exchange slabs --exchange-slabs 2 --no-e2e 2>&1 | show
echo "== gcn $N GPUs auto"; $TR benchmarks/gcn_epoch.py --epochs 8 --warmup 3 2>&1 | grep -E '^\{' | tail -1 | cut -c1-1300
echo "== gcn $N GPUs gather"; $TR benchmarks/gcn_epoch.py --epochs 8 --warmup 3 --schedule gather 2>&1 | grep -E '^\{' | tail -1 | cut -c1-1300
