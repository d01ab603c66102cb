S="python benchmarks/sweep_kernel.py"
$S --shape products --set occupancy3=0 --set occupancy3=1 --set occupancy3=1,chunk=4096 --set occupancy3=1,chunk=16384 2>&1 | grep '^{' | cut -c1-300
$S --shape products --rows-frac 0.125 --set occupancy3=0 --set occupancy3=1 2>&1 | grep '^{' | cut -c1-300
$S --shape products --dim 48 --set occupancy3=0 --set occupancy3=1 2>&1 | grep '^{' | cut -c1-300
$S --shape products --dim 256 --set occupancy3=0 --set occupancy3=1 2>&1 | grep '^{' | cut -c1-300
