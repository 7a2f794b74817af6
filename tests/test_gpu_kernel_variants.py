"""The specialised kernels of the HDR hot path (k_pixel_hdr, k_pixel_planes, k_polar_up2 MODE 2)
must produce the bytes of the general kernels they replace. The general kernels are selected by
environment switches read once per process (PLB_PIXEL_HDR / PLB_PIXEL_PLANES / PLB_UP2_HDR = 0),
so the reference render runs in a child process; sizes that are not multiples of the 64x32 /
32x32 tiles exercise the edge handling of the x-fastest tails."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [
    # (w, h, target fmt, depth, kwargs)
    (256, 144, "rgba8", 8, dict(upscaler="ewa_lanczos", deband=True, peak=2, dither=1, cmap_hq=True)),
    (250, 142, "rgba16", 16, dict(upscaler="ewa_lanczos", deband=True, peak=2, cmap_hq=True, frame_index=5)),
    (192, 108, "rgba16", 16, dict(upscaler="ewa_lanczos", peak=2, cmap_hq=True)),
    (256, 144, "rgba8", 8, dict(upscaler="ewa_lanczos", deband=True, peak=2, dither=1, cmap_hq=True,
                                grain=(1234, 1), frame_index=3)),
]


def render_all():
    import gpu_helpers as G
    import pipeline_oracle as P
    import plbtest as T
    import test_gpu_render as R
    outs = []
    with G.Gpu() as g:
        for i, (w, h, fmt, depth, kw) in enumerate(CASES):
            g.lib.shim_renderer_reset(__import__("ctypes").c_void_p(g.ctx))
            planes = T.yuv420p10(w, h, seed=20 + i, kind="noise" if i == 2 else "rings")
            img = P.Frame(planes, T.SYS["BT_2020_NC"], csp=R.HDR10)
            tgt = P.Target(w, h, fmt, sample_depth=depth, color_depth=depth, csp=R.SDR)
            got, peak = R.gpu_render(g, img, tgt, **kw)
            outs.append(got)
            outs.append(np.asarray(peak, np.float32))
    return outs


def test_specialised_kernels_match_general(tmp_path):
    out = str(tmp_path / "general.npz")
    env = dict(os.environ, PLB_PIXEL_HDR="0", PLB_PIXEL_PLANES="0", PLB_UP2_HDR="0")
    r = subprocess.run([sys.executable, os.path.abspath(__file__), out], env=env, cwd=HERE,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    general = np.load(out)
    for k in ("PLB_PIXEL_HDR", "PLB_PIXEL_PLANES", "PLB_UP2_HDR"):
        assert os.environ.get(k, "") in ("", "5", "4", "1"), f"{k} is set in the test environment"
    special = render_all()
    assert len(special) == len(general.files)
    for i, a in enumerate(special):
        b = general[f"arr_{i}"]
        assert a.shape == b.shape and a.dtype == b.dtype
        assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), f"output {i} differs"


if __name__ == "__main__":
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    np.savez(sys.argv[1], *render_all())
