#!/bin/bash
# round 2, 1 GPU: full GPU suite (no -x), bench, GCN / GIN epochs with the row-sorted copy
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2_c18_tests_all.log 2>&1; echo "all tests rc=$?"; tail -8 gpurun_out/r2_c18_tests_all.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_c18_bench_1.json 2> gpurun_out/r2_c18_bench_1.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/r2_c18_bench_1.json").read().strip().splitlines()[-1])
    print("bench1", round(d["ms_per_step"],4), d["value"], d["roofline"], d["parity"], d["e2e"]["ms_per_step"], d["cpu_baseline"])
    for s in d["extra"]["sweep"]:
        print("  ", s["config"]["workload"][:40], s["config"]["dim"], s["config"]["classifier"], round(s["ms_per_step"],4), s["roofline"]["bound"], round(s["roofline"]["frac"],3), s["parity"].get("cpu_rel_fro"), s["config"]["preprocess_ms"])
except Exception as e:
    print("no line", e); print(open("gpurun_out/r2_c18_bench_1.err").read()[-2000:])
PY
timeout 300 python benchmarks/gcn_epoch.py --epochs 10 --warmup 3 > gpurun_out/r2_c18_gcn_1.json 2> gpurun_out/r2_c18_gcn_1.err; echo "gcn1 rc=$?"; cut -c 1-700 gpurun_out/r2_c18_gcn_1.json
timeout 300 python benchmarks/gcn_epoch.py --shape proteins --model gin --feat 256 --hidden 256 --classes 112 --classifier b200 --dense --epochs 10 --warmup 3 > gpurun_out/r2_c18_gin_1.json 2> gpurun_out/r2_c18_gin_1.err; echo "gin1 rc=$?"; cut -c 1-500 gpurun_out/r2_c18_gin_1.json
