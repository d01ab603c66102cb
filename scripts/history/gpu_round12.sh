# 1 GPU: launch list (our kernels only), format sweep (bounded), GIN/GCN epochs on the proteins shape
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'hcspmm|spmm_|merge_path|preprocess' -c 40 --csv --log-file gpurun_out/launches_r1b.csv $CMD > gpurun_out/ncu_list.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_r1b.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[-14:]: print(f"{float(r[-1])/1e3:9.1f} us  {r[4][:90]}  grid {r[8]}")
PY
echo "== GIN proteins hidden 256 (shipped, all CUDA-core)"; python benchmarks/gcn_epoch.py --shape proteins --model gin --feat 256 --hidden 256 --classes 112 --epochs 8 --warmup 3 2>/dev/null | tail -1 | cut -c1-900
echo "== GIN proteins hidden 256 (b200 selector + tcgen05 dense)"; python benchmarks/gcn_epoch.py --shape proteins --model gin --feat 256 --hidden 256 --classes 112 --classifier b200 --dense --epochs 8 --warmup 3 2>/dev/null | tail -1 | cut -c1-900
echo "== format sweep"; timeout 600 python benchmarks/format_sweep.py --nnz 1000000 10000000 100000000 --deg 16 128 --dims 32 128 512 --bands 32 2048 2>&1 | grep '^{' > gpurun_out/format_sweep_r1.jsonl; wc -l gpurun_out/format_sweep_r1.jsonl
python - <<PY
import json
for l in open('gpurun_out/format_sweep_r1.jsonl'):
    d=json.loads(l)
    ks=[k for k in d if isinstance(d[k],dict)]
    print('%-8s n=%-8d deg=%-5.0f dim=%-3d '%(d['graph'],d['nodes'],d['avg_degree'],d['dim'])+'  '.join('%s %.3f(%d tc,%d dg)'%(k,d[k]['ms'],d[k]['tc_windows'],d[k]['dense_groups']) for k in ks))
PY
