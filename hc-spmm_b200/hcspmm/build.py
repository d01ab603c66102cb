"""In-tree build of the two native artefacts (sm_100a only):

  hc-spmm_b200/lib/libhcspmm.so   CUDA kernels + C ABI (include/hcspmm.h), nvcc, no torch
  hc-spmm_b200/HCSPMM.so          torch extension module `HCSPMM` (csrc/torch_shim.cpp), g++

Both are git-ignored but travel to the GPU box with the gpurun snapshot.  nvcc cross-compiles
without a GPU, so `build()` runs anywhere the toolchain is present.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import sysconfig

PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # hc-spmm_b200/
REPO_ROOT = os.path.dirname(PKG_ROOT)
CSRC = os.path.join(PKG_ROOT, "csrc")
LIB_DIR = os.path.join(PKG_ROOT, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libhcspmm.so")
EXT_PATH = os.path.join(PKG_ROOT, "HCSPMM.so")

CU_SOURCES = ["capi.cu", "preprocess.cu", "spmm.cu", "gemm.cu", "loa.cu", "umma_gemm.cu", "update_gemm.cu", "dense.cu", "dense_tma.cu", "peer.cu", "microbench.cu", "hotcols.cu", "rowsort.cu"]
HEADERS = ["common.cuh", "umma.cuh", os.path.join(REPO_ROOT, "include", "hcspmm.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps)


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libhcspmm.so")


def build_lib(force: bool = False, verbose: bool = False) -> str:
    """One nvcc process per translation unit (objects under lib/obj/, rebuilt only when the source or a
    header is newer), run in parallel, then one link."""
    from concurrent.futures import ThreadPoolExecutor
    srcs = [os.path.join(CSRC, s) for s in CU_SOURCES]
    hdrs = [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    if not force and _newer(LIB_PATH, srcs + hdrs):
        return LIB_PATH
    obj_dir = os.path.join(LIB_DIR, "obj")
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = _nvcc()
    flags = [f for f in NVCC_FLAGS if f != "-shared"]

    def compile_one(src):
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        if not force and _newer(obj, [src] + hdrs):
            return obj
        cmd = [nvcc, *flags, "-ccbin", "g++", "-c", "-o", obj, src]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd))
        subprocess.check_call(cmd)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_one, srcs))
    subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "g++",
                           "-o", LIB_PATH, *objs])
    return LIB_PATH


def build_ext(force: bool = False, verbose: bool = False) -> str:
    """Compile the torch shim with g++ directly (one translation unit, no CUDA code)."""
    src = os.path.join(CSRC, "torch_shim.cpp")
    deps = [src, os.path.join(REPO_ROOT, "include", "hcspmm.h"), LIB_PATH]
    if not force and _newer(EXT_PATH, deps):
        return EXT_PATH
    import torch
    from torch.utils import cpp_extension as ce

    inc = [f"-I{p}" for p in ce.include_paths()] + [f"-I{sysconfig.get_paths()['include']}"]
    cuda_home = ce.CUDA_HOME or "/usr/local/cuda"
    inc.append(f"-I{os.path.join(cuda_home, 'include')}")
    torch_lib = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = ["g++", "-O2", "-fPIC", "-shared", "-std=c++17", "-w",
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
           "-DTORCH_EXTENSION_NAME=HCSPMM", "-DTORCH_API_INCLUDE_EXTENSION_H",
           *inc, src, "-o", EXT_PATH,
           f"-L{LIB_DIR}", "-lhcspmm", f"-L{torch_lib}", "-lc10", "-lc10_cuda", "-ltorch_cpu",
           "-ltorch_cuda", "-ltorch", "-ltorch_python",
           "-Wl,-rpath,$ORIGIN/lib", f"-Wl,-rpath,{torch_lib}"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return EXT_PATH


def build_all(force: bool = False, verbose: bool = False):
    return build_lib(force, verbose), build_ext(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
