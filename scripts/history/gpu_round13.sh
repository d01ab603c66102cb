python -m pytest tests/test_gpu_parity.py tests/test_gpu_module.py tests/test_gpu_peer.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python benchmarks/format_sweep.py --nnz 10000000 --deg 16 128 --dims 32 128 512 --bands 32 2>&1 | grep '^{' > gpurun_out/format_sweep_b.jsonl
python - <<PY
import json
for l in open('gpurun_out/format_sweep_b.jsonl'):
    d=json.loads(l)
    ks=[k for k in d if isinstance(d[k],dict)]
    print('%-8s n=%-8d deg=%-5.0f dim=%-3d '%(d['graph'],d['nodes'],d['avg_degree'],d['dim'])+'  '.join('%s %.3f(%d tc,%d dg)'%(k,d[k]['ms'],d[k]['tc_windows'],d[k]['dense_groups']) for k in ks))
PY
