S="python benchmarks/sweep_kernel.py"
$S --shape products --dim 100 --set occupancy3=0 --set occupancy3=1 2>&1 | grep '^{' | cut -c1-300
$S --shape products --dim 104 --set occupancy3=1 2>&1 | grep '^{' | cut -c1-300
$S --shape products --dim 36 --set occupancy3=0 --set occupancy3=1 2>&1 | grep '^{' | cut -c1-300
python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fp32 or balanced or strided" 2>&1 | tail -2
