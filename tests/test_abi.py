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


def _header_enum(name, prefix):
    import re
    src = open(os.path.join(T.ROOT, "include", "libplacebo_b200.h")).read()
    body = re.search(r"enum %s\s*\{(.*?)\};" % name, src, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    out, val = {}, -1
    for item in body.split(","):
        item = item.strip()
        if not item:
            continue
        if "=" in item:
            k, v = (x.strip() for x in item.split("="))
            val = int(v, 0)
        else:
            k, val = item, val + 1
        out[k.replace(prefix, "")] = val
    return out


def test_python_enums_match_header():
    """The Python front-ends carry copies of the C enums; they must not drift."""
    from libplacebo_b200 import stream as S
    sys_h = _header_enum("pl_color_system", "PL_COLOR_SYSTEM_")
    prim_h = _header_enum("pl_color_primaries", "PL_COLOR_PRIM_")
    trc_h = _header_enum("pl_color_transfer", "PL_COLOR_TRC_")
    for table, hdr in ((S.SYS, sys_h), (S.PRIM, prim_h), (S.TRC, trc_h),
                       (T.SYS, sys_h), (T.PRIM, prim_h), (T.TRC, trc_h)):
        for k, v in table.items():
            assert hdr[k] == v, (k, v, hdr[k])


def test_forwarding_headers_compile(tmp_path):
    """<libplacebo/*.h> forwarding headers: a translation unit written against the reference's
    include paths compiles against this repository's include directory."""
    import subprocess
    src = tmp_path / "t.c"
    src.write_text("#include <libplacebo/log.h>\n#include <libplacebo/gpu.h>\n#include <libplacebo/cuda.h>\n"
                   "#include <libplacebo/renderer.h>\n#include <libplacebo/dispatch.h>\n"
                   "#include <libplacebo/shaders/sampling.h>\n#include <libplacebo/shaders/colorspace.h>\n"
                   "#include <libplacebo/shaders/dithering.h>\n#include <libplacebo/shaders/film_grain.h>\n"
                   "int f(void) { struct pl_render_params p = pl_render_default_params; return !!p.upscaler; }\n")
    subprocess.run(["gcc", "-std=gnu11", "-Wall", "-Werror", "-I", os.path.join(T.ROOT, "include"), "-c", str(src),
                    "-o", str(tmp_path / "t.o")], check=True)


def test_struct_layouts_match_reference_headers():
    """every parameter / state struct that crosses the boundary has the same
    size and member offsets as in the reference's public headers, so that a caller built against
    <libplacebo/*.h> can be relinked without recompiling. The same source (oracle/t0_flat.inc,
    abi_layout) is compiled once against the reference headers and once against
    include/libplacebo_b200.h; 46 structs / 335 members."""
    import ctypes as C
    harvest = os.path.join(T.ROOT, "oracle", "_ref", "libplref_harvest.so")
    if not os.path.exists(harvest):
        import pytest
        pytest.skip("reference not built")

    def dump(lib, prefix):
        buf = C.create_string_buffer(1 << 16)
        n = getattr(lib, prefix + "abi_layout")(buf, len(buf))
        assert 0 < n < len(buf)
        return buf.value.decode().splitlines()

    ref, mine = dump(C.CDLL(harvest), "hvt_"), dump(T.shim(), "plt_")
    assert len(ref) == len(mine) > 300
    diff = [(a, b) for a, b in zip(ref, mine) if a != b]
    assert not diff, diff[:10]
