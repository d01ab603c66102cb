"""hcspmm -- host side of the B200-native HC-SpMM hot path.

  hcspmm.capi     ctypes binding of the C ABI (include/hcspmm.h), raw device pointers
  hcspmm.graphs   seeded synthetic graphs in the reference's CSR convention
  hcspmm.build    in-tree build of libhcspmm.so and the `HCSPMM` torch extension

The drop-in module for the reference's ``import HCSPMM`` is ``hc-spmm_b200/HCSPMM.so`` (put
``hc-spmm_b200`` on sys.path).  Nothing here falls back to the CPU.
"""
BLK_H = 16   # reference hybrid_kernel/config.h:4, config.py:1
BLK_W = 8    # reference hybrid_kernel/config.h:5, config.py:2
WARP_SIZE = 32
