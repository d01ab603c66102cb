#!/bin/bash
# round 2, 4 GPUs, final build: bench (Reddit headline + products sub-record), GCN / GIN epochs
mkdir -p gpurun_out
run() { n=$1; name=$2; to=$3; shift; shift; shift
  timeout $to python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 "$@" 2> gpurun_out/r2_c23_$name.err | grep '^{' > gpurun_out/r2_c23_$name.json
  echo "$name rc=$? $(head -c 200 gpurun_out/r2_c23_$name.json)"; }
run 4 bench_4 240 bench.py --gpus 4 --steps 20 --warmup 5
run 4 gcn_4 200 benchmarks/gcn_epoch.py --epochs 10 --warmup 3
run 4 gin_4 200 benchmarks/gcn_epoch.py --shape proteins --model gin --feat 256 --hidden 256 --classes 112 --classifier b200 --dense --epochs 10 --warmup 3
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_c23_bench_4.json").read())
print("bench4", round(d["ms_per_step"],4), d["value"], d["config"]["phases"], d["parity"]["rel_fro"])
p=d["extra"]["products"]; print("  products", round(p["ms_per_step"],4), p["phases"], p["parity"]["rel_fro"])
for f in ("gcn_4","gin_4"):
    g=json.loads(open(f"gpurun_out/r2_c23_{f}.json").read()); print(f, g["value"], g["phases"], g["loss_vs_single_gpu"]["max_rel_diff"])
PY
