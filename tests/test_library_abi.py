"""CPU tests of the boundary: the C-ABI library loads and exports every symbol include/hcspmm.h
declares; argument errors are reported without a GPU; the torch extension imports and exposes
the reference's 18 names.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "hcspmm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hcspmm_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    from hcspmm import capi
    assert declared_symbols() == sorted(capi.EXPORTS)


def test_library_exports_every_declared_symbol():
    from hcspmm import build, capi
    assert os.path.exists(build.LIB_PATH), "libhcspmm.so not built: run __graft_entry__.build()"
    L = ctypes.CDLL(build.LIB_PATH)
    for sym in declared_symbols():
        assert hasattr(L, sym), sym
    assert capi.lib().hcspmm_version() >= 100


def test_argument_errors_need_no_gpu():
    from hcspmm import capi
    L = capi.lib()
    rc = L.hcspmm_preprocess(None, None, 32, 0, 5, 0, None, None, None, None, None, 0, None)
    assert rc == -1 and b"n_windows" in L.hcspmm_last_error()
    rc = L.hcspmm_spmm(None, 4, 4, None, None, None, None, None, None, 4, 0, 4, 0, 0, None, 4, None)
    assert rc == -1
    rc = L.hcspmm_spmm(None, 4, 4, None, None, None, None, None, None, 4, 0, 4, 9, 0, None, 4, None)
    assert rc == -1
    assert L.hcspmm_set_tuning(b"nope", 1) == -1
    old = L.hcspmm_set_tuning(b"slab", 64)
    assert L.hcspmm_set_tuning(b"slab", old) == 64
    assert L.hcspmm_preprocess_workspace_bytes(1600, 10) >= 100 * 4


def test_no_cpu_fallback_in_product():
    """The product path must not import the oracle or compute on the CPU."""
    pkg = os.path.join(ROOT, "hc-spmm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "liboracle" not in text, f


def test_torch_extension_imports_and_lists_reference_names():
    import torch  # noqa: F401
    import HCSPMM
    for name in ["preprocess", "forward", "forward_more", "forward_fixed32", "forward_fixed32_fused",
                 "forward_final_fused", "forward_fixed64", "forward_fixed64_fused", "forward_final_fused_64",
                 "forward_GIN_final_fused", "backward", "backward_fixed32", "backward_fixed32_fused",
                 "backward_final_fused", "backward_fixed64", "backward_fixed64_fused", "backward_final_fused_64",
                 "backward_GIN_final_fused"]:
        assert hasattr(HCSPMM, name), name
    with pytest.raises(RuntimeError):
        HCSPMM.forward(*[torch.zeros(1)] * 9)  # CPU tensors are rejected, no fallback
