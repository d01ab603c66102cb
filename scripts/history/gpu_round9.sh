python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3
S="python benchmarks/sweep_kernel.py"
$S --shape reddit --set balance=0 --set chunk=4096 --set chunk=8192 --set chunk=2048 --set chunk=16384 --set chunk=4096,warp_split=0 --set chunk=8192,long_row=100000 2>&1 | grep '^{' | cut -c1-400
$S --shape reddit --rows-frac 0.125 --set balance=0 --set chunk=4096 --set chunk=2048 --set chunk=8192 --set chunk=4096,warp_split=0 2>&1 | grep '^{' | cut -c1-400
$S --shape products --set balance=0 --set chunk=4096 --set chunk=8192 --set chunk=2048 --set chunk=8192,warp_split=0 2>&1 | grep '^{' | cut -c1-400
$S --shape products --rows-frac 0.125 --set balance=0 --set chunk=4096 --set chunk=2048 --set chunk=8192 2>&1 | grep '^{' | cut -c1-400
$S --shape envelope --set balance=0 --set chunk=4096 --set chunk=8192 --set chunk=16384 2>&1 | grep '^{' | cut -c1-400
$S --shape proteins --set balance=0 --set chunk=4096 --set chunk=8192 --set chunk=8192,warp_split=0 2>&1 | grep '^{' | cut -c1-400
$S --shape reddit --dim 64 --set balance=0 --set chunk=4096 --set chunk=8192 2>&1 | grep '^{' | cut -c1-400
$S --shape reddit --dim 512 --set balance=0 --set chunk=4096 --set chunk=8192 2>&1 | grep '^{' | cut -c1-400
$S --shape reddit --precision bf16 --set balance=0 --set chunk=4096 --set chunk=8192 2>&1 | grep '^{' | cut -c1-400
