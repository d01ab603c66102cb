N=${N:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR scripts/gpu_peer_check.py 2>&1 | grep -vE "Warning|warn|^\s*$|\*\*\*" | tail -6
run() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.log 2>&1; grep -E '^\{' gpurun_out/$name.log | tail -1 > gpurun_out/$name.json; if [ -s gpurun_out/$name.json ]; then python - <<PY
import json
d=json.load(open('gpurun_out/$name.json'))
c=d['config']
print('$name: value %.2f %s  ms/step %s  e2e %s  phases %s  schedule %s' % (d['value'], d['unit'], d.get('ms_per_step'), (d.get('e2e') or {}).get('value'), c.get('phases') or d.get('phases'), c.get('exchange') or c.get('schedule')))
PY
else grep -vE "Warning|warn" gpurun_out/$name.log | tail -15; fi; }
run r1c_scale_bench_reddit_$N $TR bench.py --gpus $N --steps 10 --warmup 3
run r1c_scale_bench_products_$N $TR bench.py --gpus $N --steps 10 --warmup 3 --shape products --no-e2e
run r1c_scale_gcn_products_$N $TR benchmarks/gcn_epoch.py --epochs 8 --warmup 3
run r1c_scale_bench_reddit_gather_$N $TR bench.py --gpus $N --steps 10 --warmup 3 --exchange gather --no-e2e
