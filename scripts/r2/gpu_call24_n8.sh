#!/bin/bash
# round 2, 8 GPUs, final build: bench (Reddit headline + products sub-record), products sweep (in-place x row blocks),
# GCN / GIN epochs, BF16 operand
mkdir -p gpurun_out
run() { n=$1; name=$2; to=$3; shift; shift; shift
  timeout $to python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 "$@" 2> gpurun_out/r2_c24_$name.err | grep '^{' > gpurun_out/r2_c24_$name.json
  echo "$name rc=$? $(head -c 200 gpurun_out/r2_c24_$name.json)"; }
run 8 bench_8 240 bench.py --gpus 8 --steps 20 --warmup 5
run 8 sweep_8 300 scripts/r2/inplace_sweep.py --refs 0 1 --row-blocks 1 2 --steps 20
cat gpurun_out/r2_c24_sweep_8.json | cut -c1-330
run 8 gcn_8 200 benchmarks/gcn_epoch.py --epochs 10 --warmup 3
run 8 gin_8 200 benchmarks/gcn_epoch.py --shape proteins --model gin --feat 256 --hidden 256 --classes 112 --classifier b200 --dense --epochs 10 --warmup 3
run 8 bench_8_bf16 200 bench.py --gpus 8 --steps 20 --warmup 5 --precision bf16 --no-e2e --no-cpu-baseline
python - <<'PY'
import json
for f in ("bench_8","bench_8_bf16"):
    d=json.loads(open(f"gpurun_out/r2_c24_{f}.json").read())
    print(f, round(d["ms_per_step"],4), d["value"], d["config"]["phases"], d["parity"]["rel_fro"], (d.get("e2e") or {}).get("ms_per_step"))
    p=d["extra"]["products"]; print("  products", round(p["ms_per_step"],4), p["phases"], p["parity"]["rel_fro"])
for f in ("gcn_8","gin_8"):
    g=json.loads(open(f"gpurun_out/r2_c24_{f}.json").read()); print(f, g["value"], g["phases"], g["loss_vs_single_gpu"]["max_rel_diff"])
PY
