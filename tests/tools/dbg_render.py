"""Developer tool: `--run` renders the C3-like cases on the GPU box and stores outputs + peak
state in gpurun_out/dbg_render.npz; without arguments compares with the pipeline oracle here."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import pipeline_oracle as P  # noqa: E402
import plbtest as T  # noqa: E402
import test_gpu_render as R  # noqa: E402

OUT = os.path.join(os.path.dirname(__file__), "..", "..", "gpurun_out", "dbg_render.npz")
SDR, HDR10 = R.SDR, R.HDR10


def cases():
    c = {}
    for name, kw in [
        ("plain", dict()),
        ("peak", dict(peak=2)),
        ("peak_lq", dict(peak=1)),
        ("ewa", dict(upscaler="ewa_lanczos")),
        ("ewa_peak", dict(upscaler="ewa_lanczos", peak=2, cmap_hq=True)),
        ("deband", dict(deband=True)),
        ("all", dict(upscaler="ewa_lanczos", deband=True, peak=2, cmap_hq=True)),
    ]:
        for kind in ("rings", "noise"):
            c[f"{name}_{kind}"] = (kind, kw)
    return c


def frame(kind):
    planes = T.yuv420p10(256, 144, seed=8, kind=kind)
    img = P.Frame(planes, T.SYS["BT_2020_NC"], csp=HDR10)
    tgt = P.Target(256, 144, "rgba16", sample_depth=16, color_depth=16, csp=SDR)
    return img, tgt


def run():
    import ctypes as C
    import gpu_helpers
    out = {}
    with gpu_helpers.Gpu() as gpu:
        for name, (kind, kw) in cases().items():
            gpu.lib.shim_renderer_reset(C.c_void_p(gpu.ctx))
            img, tgt = frame(kind)
            got, peak = R.gpu_render(gpu, img, tgt, **kw)
            out[name] = got
            out[name + "__peak"] = peak
    np.savez_compressed(OUT, **out)
    print("saved", len(out))


def check():
    got = np.load(OUT)
    for name, (kind, kw) in cases().items():
        img, tgt = frame(kind)
        want, ps = P.render(img, tgt, **kw)
        g = got[name]
        d = np.abs(g[..., :3].astype(np.int32) - want[..., :3].astype(np.int32))
        print(f"{name:18s} max {d.max():5d} frac>0 {np.mean(d > 0):.4f} frac>1 {np.mean(d > 1):.4f} "
              f"peak got {got[name + '__peak']} want {ps}")


if __name__ == "__main__":
    run() if "--run" in sys.argv else check()
