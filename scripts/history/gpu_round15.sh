S="python benchmarks/sweep_kernel.py"
$S --shape products --set occupancy3=1 --set occupancy3=2 --set occupancy3=2,chunk=4096 2>&1 | grep '^{' | cut -c1-300
$S --shape products --rows-frac 0.125 --set occupancy3=1 --set occupancy3=2 2>&1 | grep '^{' | cut -c1-300
$S --shape products --dim 48 --set occupancy3=1 --set occupancy3=2 2>&1 | grep '^{' | cut -c1-300
$S --shape products --dim 256 --set occupancy3=1 --set occupancy3=2 2>&1 | grep '^{' | cut -c1-300
