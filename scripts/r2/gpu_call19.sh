#!/bin/bash
# round 2, 1 GPU: A/B of the row-sorted copy in the bench and in the GCN epoch (same box, same process order)
mkdir -p gpurun_out
for v in "sort:" "nosort:--no-row-sort" "sort2:" "nosort2:--no-row-sort"; do
  name=${v%%:*}; flags=${v#*:}
  timeout 300 python bench.py --shape products --steps 20 --warmup 5 --no-extra --no-e2e --no-cpu-baseline $flags > gpurun_out/r2_c19_bench_products_$name.json 2> gpurun_out/r2_c19_$name.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r2_c19_bench_products_$name.json").read().strip().splitlines()[-1])
print("$name", round(d["ms_per_step"],4), d["step_ms"], d["config"]["preprocess_ms"], d["clocks"])
PY
done
timeout 300 python benchmarks/gcn_epoch.py --epochs 10 --warmup 3 --no-row-sort > gpurun_out/r2_c19_gcn_1_nosort.json 2> gpurun_out/r2_c19_gcn_1_nosort.err; echo "gcn1 nosort rc=$?"; cut -c 1-120 gpurun_out/r2_c19_gcn_1_nosort.json
python - <<'PY'
import json
for f in ("gpurun_out/r2_c19_gcn_1_nosort.json",):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["value"], d["phases"])
PY
