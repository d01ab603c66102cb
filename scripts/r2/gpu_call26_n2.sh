#!/bin/bash
# round 2, 2 GPUs, final build: bench (Reddit headline + products sub-record), GCN / GIN epochs
mkdir -p gpurun_out
run() { n=$1; name=$2; to=$3; shift; shift; shift
  timeout $to python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 "$@" 2> gpurun_out/r2_c26_$name.err | grep '^{' > gpurun_out/r2_c26_$name.json
  echo "$name rc=$? $(head -c 200 gpurun_out/r2_c26_$name.json)"; }
run 2 bench_2 240 bench.py --gpus 2 --steps 20 --warmup 5
run 2 gcn_2 200 benchmarks/gcn_epoch.py --epochs 10 --warmup 3
run 2 gin_2 200 benchmarks/gcn_epoch.py --shape proteins --model gin --feat 256 --hidden 256 --classes 112 --classifier b200 --dense --epochs 10 --warmup 3
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_c26_bench_2.json").read())
print("bench2", round(d["ms_per_step"],4), d["value"], d["config"]["phases"], d["parity"]["rel_fro"])
p=d["extra"]["products"]; print("  products", round(p["ms_per_step"],4), p["phases"], p["parity"]["rel_fro"])
for f in ("gcn_2","gin_2"):
    g=json.loads(open(f"gpurun_out/r2_c26_{f}.json").read()); print(f, g["value"], g["phases"], g["loss_vs_single_gpu"]["max_rel_diff"])
PY
