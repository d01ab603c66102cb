#!/bin/bash
# round 2, 8 GPUs, final build: config-5 sweep at 500 M entries again (the first run had the in-place auto rule of that
# build on the all-CUDA selectors), bench with rows referenced once read in place
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 benchmarks/format_sweep.py --nnz 500000000 --deg 128 --dims 32 256 --bands 256 --skip-all-tc 2> gpurun_out/r2_c33_sweep.err | grep '^{' > gpurun_out/r2_format_sweep_8gpu_final.jsonl; echo "sweep rc=$?"; tail -2 gpurun_out/r2_c33_sweep.err; cut -c 1-300 gpurun_out/r2_format_sweep_8gpu_final.jsonl
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 --direct-refs 1 --no-e2e --no-cpu-baseline 2> gpurun_out/r2_c33_bench_8_t1.err | grep '^{' > gpurun_out/r2_c33_bench_8_t1.json; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_c33_bench_8_t1.json").read())
print("bench8 T1", round(d["ms_per_step"],4), d["config"]["phases"], d["parity"]["rel_fro"])
p=d["extra"]["products"]; print("  products", round(p["ms_per_step"],4), p["phases"], p["parity"]["rel_fro"])
PY
