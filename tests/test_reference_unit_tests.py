"""The reference's OWN unit tests for the host code of the path (src/tests/filters.c,
tone_mapping.c) compiled in place from /root/reference — nothing is copied — against
the reference's public headers and linked against libplacebo_b200.so instead of libplacebo:
the drop-in claim of INTEGRATION.md, checked with the maintainers' own assertions (filter LUT
accuracy and row sums, tone-curve monotonicity / clip identity / PQ constants). The reference's
colorspace.c / common.c unit tests exercise API surface outside SURVEY.md section 8 (names, crop
helpers, CPU transfer functions, cone matrices), which this library no longer carries.
Skipped where the reference tree is absent (the GPU box)."""
import os
import subprocess

import pytest

import plbtest as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("PL_REFERENCE", "/root/reference")
REFBUILD = os.path.join(ROOT, "oracle", "_ref")

pytestmark = pytest.mark.skipif(
    not (os.path.isdir(os.path.join(REF, "src", "tests")) and os.path.isdir(os.path.join(REFBUILD, "inc"))),
    reason="reference tree / oracle/_ref headers not present")


@pytest.mark.parametrize("name", ["filters", "tone_mapping"])
def test_reference_unit_test_passes_against_product_library(name, tmp_path):
    T.product_t0()  # makes sure libplacebo_b200.so is built
    exe = str(tmp_path / name)
    libdir = os.path.join(ROOT, "libplacebo_b200")
    cmd = ["gcc", "-std=c11", "-O1", "-D_GNU_SOURCE", "-DPL_HAVE_PTHREAD", "-DPTHREAD_HAS_SETCLOCK",
           "-DPL_HAVE_DOVI", "-w",
           # -fPIC: extern data (pl_filter_function_gaussian, ...) through the GOT, so that addresses
           # compare equal with the library's own (it is linked -Bsymbolic; no copy relocations)
           "-fPIC",
           "-I" + REFBUILD, "-I" + os.path.join(REFBUILD, "inc"),
           "-I" + os.path.join(REF, "src", "include"), "-I" + os.path.join(REF, "src"),
           "-I" + os.path.join(REF, "src", "tests"),
           "-I" + os.path.join(REF, "3rdparty", "fast_float", "include"),
           os.path.join(REF, "src", "tests", name + ".c"),
           os.path.join(ROOT, "tests", "c", "ref_tests_shim.c"),
           "-o", exe, "-L" + libdir, "-lplacebo_b200", "-Wl,-rpath," + libdir, "-lm", "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, "does not link against the product library:\n" + r.stderr[-3000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-1500:] + "\n" + r.stderr[-3000:])
