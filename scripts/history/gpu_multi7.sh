N=${N:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 300 $TR scripts/gpu_peer_check.py 2>&1 | grep -E "rank [0-9]+ slabs|PEER|Error|error" | grep -v Warning | head -12
run() { name=$1; shift; timeout 300 "$@" > gpurun_out/$name.log 2>&1; grep -E '^\{' gpurun_out/$name.log | tail -1 > gpurun_out/$name.json; if [ -s gpurun_out/$name.json ]; then python - <<PY
import json
d=json.load(open('gpurun_out/$name.json'))
c=d['config']
print('$name: value %.2f %s  ms/step %s  phases %s  %s' % (d['value'], d['unit'], d.get('ms_per_step'), c.get('phases') or d.get('phases'), (c.get('exchange') or c.get('schedule'))[-40:]))
PY
else grep -vE "Warning|warn" gpurun_out/$name.log | tail -15; fi; }
run r1e_reddit_p1_$N $TR bench.py --gpus $N --steps 10 --warmup 3 --no-e2e
run r1e_reddit_p2_$N $TR bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --exchange-passes 2
run r1e_reddit_p2c16_$N $TR bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --exchange-passes 2 --overlap-ctas 16
run r1e_products_p1_$N $TR bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --shape products
run r1e_products_p2_$N $TR bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --shape products --exchange-passes 2
