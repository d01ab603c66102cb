python benchmarks/hcspmm_main.py --dataset example --model gcn --single_kernel --dim 32 2>&1 | grep -vE "Warning|warn" | tail -8
python benchmarks/hcspmm_main.py --dataset example_rmat --model gin --epochs 20 --num_layers 3 --hidden 64 --dim 32 2>&1 | grep -vE "Warning|warn" | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
$TR bench.py --gpus 2 --steps 5 --warmup 3 2>&1 | grep -E '^\{' | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=2 value', d['value'], 'ms', d['ms_per_step'], 'roofline', d['roofline'], 'launches', d['gpu_launches'])"
