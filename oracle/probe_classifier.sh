#!/bin/sh
# Prints the PTX nvcc 12.9 emits for the reference's selector expressions
# (hybrid_all_kernel.cu:261 intended, :262 shipped), compiled the way the reference is compiled
# (default -fmad=true).  oracle/hcspmm_oracle.c and csrc/preprocess.cu restate this sequence.
set -e
T=$(mktemp -d)
cat > $T/cls.cu <<'CU'
__global__ void k(int size, unsigned num_window_edges, int num, int* intended, int* shipped){
  intended[0] = (size > 32 || (float)size * 0.19854024 - ((float)num_window_edges / (num * 16 * 8)) * 6.578043 - 3.14922857 > 0) ? 0:1;
  shipped[0] = (float)size * 0.19854024 - ((float)num_window_edges / (num * 16 * 8)) * 6.578043 - 3.14922857 ? 0:1;
}
CU
nvcc -arch=sm_100 -ptx $T/cls.cu -o $T/cls.ptx
grep -E 'cvt|div|mul|fma|add|setp|selp' $T/cls.ptx
rm -rf $T
