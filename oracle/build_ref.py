"""Compile the UNMODIFIED reference extension into oracle/_ref/HCSPMM_ref.so.

TEST INFRASTRUCTURE ONLY.  Sources are compiled from where they lie under
/root/reference/hybrid_kernel (hybrid_all.cpp, hybrid_all_kernel.cu, config.h) -- nothing
is copied into the repo; only the module name differs (TORCH_EXTENSION_NAME=HCSPMM_ref) so
that it can be imported next to our own `HCSPMM`.  The reference's setup.py is not run; this
is the same two-file CUDAExtension it describes (setup.py:5-11), built for sm_100.

Run in the development container (the reference is not present on the GPU box; the built
.so travels there with the gpurun snapshot because oracle/_ref/ is git-ignored only).
"""
import os
import sys

REF = os.environ.get("HCSPMM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")


def build(verbose: bool = False):
    src = [os.path.join(REF, "hybrid_kernel", f) for f in ("hybrid_all.cpp", "hybrid_all_kernel.cu")]
    if not all(os.path.exists(s) for s in src):
        print("reference not present: skipping oracle/_ref/HCSPMM_ref.so")
        return None
    target = os.path.join(OUT, "HCSPMM_ref.so")
    if os.path.exists(target) and all(os.path.getmtime(s) <= os.path.getmtime(target) for s in src):
        return target
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.pop("CC", None)
    os.environ.pop("CXX", None)
    from torch.utils.cpp_extension import load
    load(name="HCSPMM_ref", sources=src, build_directory=OUT, verbose=verbose,
         extra_cflags=["-w"], extra_cuda_cflags=["-w"], is_python_module=False)
    return target


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
