"""Debug helper (GPU box): where does the C4 full-size frame differ from the oracle?"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import ctypes as C
import numpy as np
import gpu_helpers as G, pipeline_oracle as P, plbtest as T
from test_gpu_render import HDR10, SDR, gpu_render
from libplacebo_b200 import stream as S

def run(g, w, h, seed, fi, deband=True, peak=2, fmt="rgba16", depth=16):
    g.lib.shim_renderer_reset(C.c_void_p(g.ctx))
    img = P.Frame(S.synth_yuv420p10(w, h, seed=3), T.SYS["BT_2020_NC"], csp=HDR10)
    tgt = P.Target(w, h, fmt, sample_depth=depth, color_depth=depth, csp=SDR)
    kw = dict(upscaler="ewa_lanczos", deband=deband, peak=peak, dither=0, cmap_hq=True, frame_index=fi)
    got, pk = gpu_render(g, img, tgt, grain=(seed, 1), **kw)
    want, ps = P.render(img, tgt, film_grain=(seed, 1), **kw)
    d = np.abs(got[..., :3].astype(np.int64) - want[..., :3].astype(np.int64)).max(-1)
    ys, xs = np.nonzero(d)
    print(f"{w}x{h} seed={seed} fi={fi} deband={deband} peak={peak}: peak gpu={pk} oracle={ps} max={d.max()} n={len(ys)}",
          "first:", list(zip(ys[:8].tolist(), xs[:8].tolist())),
          "x mod 32:", np.bincount(xs % 32, minlength=32).tolist() if len(xs) else None,
          "y mod 32:", np.bincount(ys % 32, minlength=32).tolist() if len(ys) else None, flush=True)

with G.Gpu() as g:
    run(g, 512, 288, 7, 7)
    run(g, 512, 288, 1234, 3)
    run(g, 1920, 1080, 7, 7)
    run(g, 3840, 2160, 7, 7, peak=0)
    run(g, 3840, 2160, 7, 7, deband=False, peak=0)
    run(g, 3840, 2160, 1234, 3, peak=0)
