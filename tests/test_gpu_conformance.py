"""The reference's OWN conformance tests of the pl_gpu boundary — src/tests/gpu_tests.c
pl_buffer_tests (:10-132: static / dynamic buffers, buffer copies, endian swapping) and
pl_texture_tests (:226-291: upload -> [clear, blit] -> download round trip of every host-readable
format as 1-D, 2-D and 3-D textures, with timers and asynchronous callbacks) — compiled in place
from /root/reference against the reference's headers and run against libplacebo_b200.so on the
GPU. The executable is built here by oracle/build_conformance.sh (called from
__graft_entry__.build()) into oracle/_ref/ and travels to the GPU box; where it is absent the test
is skipped."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "gpu_conformance")


def test_reference_buffer_and_texture_tests_pass_on_the_cuda_backend():
    if not os.path.exists(EXE):
        pytest.skip("oracle/_ref/gpu_conformance not built (needs the reference tree)")
    r = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    tail = r.stdout[-2500:] + "\n" + r.stderr[-2500:]
    assert r.returncode == 0, tail
    assert "conformance: ok" in r.stdout, tail
    assert "- test endian swapping" in r.stdout and "-- 3D" in r.stdout, tail
