"""The C-ABI library loads without a GPU and exports every symbol the public header declares."""
import ctypes
import os
import re

import plbtest as T


def declared_symbols():
    src = open(os.path.join(T.ROOT, "include", "libplacebo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    funcs = set(re.findall(r"PL_API\s+[^;{]*?\b(pl_\w+)\s*\(", src))
    data = set(re.findall(r"PL_API\s+extern\s+[^;]*?\b(pl_\w+)\s*(?:\[\])?\s*;", src))
    funcs.discard("pl_log_create")          # macro-renamed to pl_log_create_365
    funcs.add("pl_log_create_365")
    return funcs, data


def test_library_exports_every_declared_symbol():
    T.product_t0()  # builds if needed
    lib = ctypes.CDLL(T.PRODUCT_SO)
    funcs, data = declared_symbols()
    assert len(funcs) > 100 and len(data) > 50
    missing = [s for s in sorted(funcs | data) if not hasattr(lib, s)]
    assert not missing, f"declared in include/libplacebo_b200.h but not exported: {missing}"


def test_product_does_not_link_the_oracle():
    import subprocess
    out = subprocess.run(["ldd", T.PRODUCT_SO], capture_output=True, text=True).stdout
    assert "oracle" not in out and "placebo_ref" not in out
    syms = subprocess.run(["nm", "-D", T.PRODUCT_SO], capture_output=True, text=True).stdout
    assert " orc_" not in syms and " hvt_" not in syms


def test_no_gpu_fails_loudly():
    """Without a CUDA device pl_cuda_create returns NULL (and logs); there is no CPU path."""
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("GPU present")
    shim = T.shim()
    shim.shim_create.restype = ctypes.c_void_p
    assert not shim.shim_create(0)
