python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
S="python benchmarks/sweep_kernel.py"
$S --shape reddit --set balance=0 --set balance=1 --set chunk=8192 2>&1 | grep '^{' | cut -c1-400
$S --shape reddit --rows-frac 0.125 --set balance=0 --set balance=1 2>&1 | grep '^{' | cut -c1-400
$S --shape products --set balance=0 --set balance=1 --set warp_split=0 --set warp_split=256 2>&1 | grep '^{' | cut -c1-400
$S --shape products --rows-frac 0.125 --set balance=0 --set balance=1 2>&1 | grep '^{' | cut -c1-400
$S --shape proteins --set balance=0 --set balance=1 --set warp_split=0 2>&1 | grep '^{' | cut -c1-400
$S --shape envelope --set balance=0 --set balance=1 2>&1 | grep '^{' | cut -c1-400
$S --shape reddit --dim 64 --set balance=0 --set balance=1 2>&1 | grep '^{' | cut -c1-400
$S --shape reddit --dim 128 --set balance=0 --set balance=1 --set chunk=4096 2>&1 | grep '^{' | cut -c1-400
echo "== gcn 1 GPU"; python benchmarks/gcn_epoch.py --epochs 8 --warmup 3 2>/dev/null | tail -1 | cut -c1-1200
