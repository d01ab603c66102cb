N=${NGPU:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
for sl in 1 2 4; do
echo "== bench $N GPUs exchange-slabs=$sl"; $TR bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --exchange-slabs $sl 2>&1 | grep -E '^\{|Error' | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['config']['phases'])"
done
