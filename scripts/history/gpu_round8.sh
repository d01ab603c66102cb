python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -5
S="python benchmarks/sweep_kernel.py"
$S --shape reddit --set balance=0 --set balance=1 --set balance=1,chunk=2048 --set balance=1,chunk=8192 --set balance=1,chunk=1024 2>&1 | grep '^{' | cut -c1-400
$S --shape reddit --rows-frac 0.125 --set balance=0 --set balance=1 --set balance=1,chunk=2048 --set balance=1,chunk=1024 2>&1 | grep '^{' | cut -c1-400
$S --shape products --set balance=0 --set balance=1 --set balance=1,chunk=2048 --set balance=1,chunk=8192 2>&1 | grep '^{' | cut -c1-400
$S --shape products --rows-frac 0.125 --set balance=0 --set balance=1 --set balance=1,chunk=2048 2>&1 | grep '^{' | cut -c1-400
$S --shape envelope --set balance=0 --set balance=1 --set balance=1,chunk=2048 --set balance=1,chunk=8192 2>&1 | grep '^{' | cut -c1-400
$S --shape proteins --set balance=0 --set balance=1 2>&1 | grep '^{' | cut -c1-400
