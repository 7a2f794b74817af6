"""pl_render_image (C ABI) against the independent CPU restatement of the reference's pass
graph (tests/pipeline_oracle.py) on identical synthetic YUV420P10 frames.
Configs mirror BASELINE.json: C2 = SDR 2x EWA-Lanczos upscale to rgba16hf, C3 = HDR10 PQ ->
BT.709 SDR with deband, chroma EWA, peak detection, tone/gamut mapping and blue-noise dither."""
import ctypes as C

import numpy as np
import pytest

import pipeline_oracle as P
import plbtest as T

pytestmark = pytest.mark.gpu


def gpu_render(gpu, img, tgt, *, upscaler="", plane_upscaler="", deband=False, peak=0,
               allow_delayed=False, dither=0, cmap_hq=False, frame_index=0, tone_fn=None,
               gamut_fn=None):
    Y, U, V = img.planes
    texs = [gpu.tex("r16", p.shape[1], p.shape[0], p) for p in (Y, U, V)]
    dst = gpu.tex(tgt.fmt, tgt.w, tgt.h)
    try:
        planes = (C.c_void_p * 3)(*texs)
        iarr = np.concatenate([[img.sys, img.levels, *img.bits], img.csp]).astype(T.f32)
        tarr = np.concatenate([[tgt.sys, tgt.levels, *tgt.bits], tgt.csp]).astype(T.f32)
        opts = np.zeros(16, np.int32)
        opts[0], opts[1], opts[2], opts[3], opts[4] = deband, peak, allow_delayed, dither, cmap_hq
        opts[5] = P.CHROMA_ENUM[img.chroma_loc]
        opts[7] = frame_index
        opts[9] = T.TONE_FNS.index(tone_fn) + 1 if tone_fn else 0
        opts[10] = T.GAMUT_FNS.index(gamut_fn) + 1 if gamut_fn else 0
        peak_out = np.zeros(2, T.f32)
        r = gpu.lib.shim_render(C.c_void_p(gpu.ctx), 3, planes, C.c_void_p(dst), T.fp(iarr),
                                T.fp(tarr), T.ip(opts), upscaler.encode(),
                                plane_upscaler.encode(), b"", T.fp(peak_out))
        assert r == 0, f"pl_render_image failed / degraded: {r}"
        return gpu.download(dst, tgt.fmt, tgt.w, tgt.h), peak_out
    finally:
        gpu.free(dst, *texs)


SDR = T.csp(T.PRIM["BT_709"], T.TRC["BT_1886"])
HDR10 = T.csp(T.PRIM["BT_2020"], T.TRC["PQ"], max_luma=1000)


@pytest.mark.parametrize("kind", ["rings", "noise"])
@pytest.mark.parametrize("loc", ["left", "center", "top_left"])
def test_c2_sdr_upscale_2x_ewa_lanczos(gpu, kind, loc):
    """Config C2 at reduced size: 4:2:0 10-bit BT.709 -> 2x EWA-Lanczos -> rgba16hf."""
    planes = T.yuv420p10(192, 108, seed=3, kind=kind)
    img = P.Frame(planes, T.SYS["BT_709"], csp=SDR, chroma_loc=loc)
    tgt = P.Target(384, 216, "rgba16hf", csp=SDR)
    got, _ = gpu_render(gpu, img, tgt, upscaler="ewa_lanczos")
    want, _ = P.render(img, tgt, upscaler="ewa_lanczos")
    ulp = T.ulp_diff_half(got, want)
    assert ulp.max() <= 1, (ulp.max(), np.mean(ulp > 0))
    assert np.mean(ulp > 0) < 1e-3


def test_c2_with_deband(gpu):
    gpu.lib.shim_renderer_reset(C.c_void_p(gpu.ctx))
    planes = T.yuv420p10(192, 108, seed=4)
    img = P.Frame(planes, T.SYS["BT_709"], csp=SDR)
    tgt = P.Target(384, 216, "rgba16hf", csp=SDR)
    got, _ = gpu_render(gpu, img, tgt, upscaler="ewa_lanczos", deband=True, frame_index=5)
    want, _ = P.render(img, tgt, upscaler="ewa_lanczos", deband=True, frame_index=5)
    ulp = T.ulp_diff_half(got, want)
    # deband decisions can flip for isolated pixels (sin/cos from different libms)
    assert np.mean(ulp > 1) < 2e-3
    assert np.median(ulp) == 0


@pytest.mark.parametrize("fmt,depth", [("rgba16", 16), ("rgba8", 8)])
def test_c3_hdr10_to_sdr(gpu, fmt, depth):
    """Config C3 at reduced size: HDR10 PQ BT.2020 -> BT.709 BT.1886, deband, chroma EWA, HQ
    peak detection (same frame), spline tone map, perceptual gamut map, blue-noise dither."""
    gpu.lib.shim_renderer_reset(C.c_void_p(gpu.ctx))
    planes = T.yuv420p10(256, 144, seed=6)
    img = P.Frame(planes, T.SYS["BT_2020_NC"], csp=HDR10)
    tgt = P.Target(256, 144, fmt, sample_depth=depth, color_depth=depth, csp=SDR)
    kw = dict(upscaler="ewa_lanczos", deband=True, peak=2, dither=1, cmap_hq=True, frame_index=0)
    got, peak = gpu_render(gpu, img, tgt, **kw)
    st = {}
    want, pstate = P.render(img, tgt, stats=st, **kw)
    # peak statistics are integers: the smoothed state must agree to float rounding
    assert abs(peak[0] - pstate[0]) < 1e-4 and abs(peak[1] - pstate[1]) < 1e-4, (peak, pstate)
    d = np.abs(got[..., :3].astype(np.int32) - want[..., :3].astype(np.int32))
    if depth == 16:
        assert np.mean(d > 1) < 5e-3, (d.max(), np.mean(d > 1))
    else:
        # dithered 8-bit: a value within float rounding of a floor() boundary may land on
        # either side
        assert np.mean(d > 0) < 5e-3 and d.max() <= 2, (d.max(), np.mean(d > 0))


def test_c3_no_deband_is_tight(gpu):
    """Without the (discontinuous) deband stage the whole chain must stay within 1 LSB of the
    16-bit output everywhere."""
    gpu.lib.shim_renderer_reset(C.c_void_p(gpu.ctx))
    planes = T.yuv420p10(256, 144, seed=8, kind="noise")
    img = P.Frame(planes, T.SYS["BT_2020_NC"], csp=HDR10)
    tgt = P.Target(256, 144, "rgba16", sample_depth=16, color_depth=16, csp=SDR)
    kw = dict(upscaler="ewa_lanczos", peak=2, cmap_hq=True)
    got, peak = gpu_render(gpu, img, tgt, **kw)
    want, pstate = P.render(img, tgt, **kw)
    d = np.abs(got[..., :3].astype(np.int32) - want[..., :3].astype(np.int32))
    assert d.max() <= 1, (d.max(), np.mean(d > 0))


def test_unsupported_features_fail_loudly(gpu):
    """pl_render_params members outside the hot path are rejected, not silently ignored."""
    import gpu_helpers
    planes = T.yuv420p10(64, 36, seed=1)
    img = P.Frame(planes, T.SYS["BT_709"], csp=SDR)
    tgt = P.Target(64, 36, "rgba16hf", csp=SDR)
    got, _ = gpu_render(gpu, img, tgt)  # baseline works
    assert np.isfinite(got.astype(np.float32)).all()
