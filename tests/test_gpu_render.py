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
               gamut_fn=None, error_diffusion=None, grain=None, contrast_recovery=0.0, luts=None):
    Y, U, V = img.planes
    keep = []
    for which, (kind, data, size) in (luts or {}).items():
        d = np.ascontiguousarray(data, np.float32)
        keep.append(d)
        gpu.lib.shim_set_lut(C.c_void_p(gpu.ctx), {"params": 0, "image": 1, "target": 2}[which],
                             {"native": 1, "normalized": 2, "conversion": 3}[kind], T.fp(d),
                             (C.c_int * 3)(*size), None, None,
                             C.c_ulonglong(hash((which, d.tobytes())) & (2 ** 63 - 1)), None)
    texs = [gpu.tex("r16", p.shape[1], p.shape[0], p) for p in (Y, U, V)]
    dst = gpu.tex(tgt.fmt, tgt.w, tgt.h)
    ptex = []
    if tgt.planes:  # planar target: shim_set_target_planes overrides `dst` for the next render
        ptex = [gpu.tex(fmt, *tgt.plane_size(i)) for i, (fmt, _, _) in enumerate(tgt.planes)]
        gpu.lib.shim_set_target_planes(C.c_void_p(gpu.ctx), len(ptex), (C.c_void_p * len(ptex))(*ptex),
                                       P.CHROMA_ENUM[tgt.chroma_loc])
    try:
        planes = (C.c_void_p * 3)(*texs)
        iarr = np.concatenate([[img.sys, img.levels, *img.bits], img.csp]).astype(T.f32)
        tarr = np.concatenate([[tgt.sys, tgt.levels, *tgt.bits], tgt.csp]).astype(T.f32)
        opts = np.zeros(16, np.int32)
        opts[0], opts[1], opts[2], opts[3], opts[4] = deband, peak, allow_delayed, dither, cmap_hq
        opts[5] = P.CHROMA_ENUM[img.chroma_loc]
        opts[7] = frame_index
        opts[8] = T.ED_KERNELS.index(error_diffusion) + 1 if error_diffusion else 0
        if grain is not None:
            if grain[1] == "h274":
                opts[12], opts[13] = 3, grain[0]
            else:
                opts[12], opts[13] = 2 if grain[1] else 1, grain[0]
        opts[15] = 1 if contrast_recovery else 0
        if getattr(img, "dovi", None) is not None:
            opts[14] = img.dovi + 1
        opts[9] = T.TONE_FNS.index(tone_fn) + 1 if tone_fn else 0
        opts[10] = T.GAMUT_FNS.index(gamut_fn) + 1 if gamut_fn else 0
        peak_out = np.zeros(2, T.f32)
        r = gpu.lib.shim_render(C.c_void_p(gpu.ctx), 3, planes, C.c_void_p(dst), T.fp(iarr),
                                T.fp(tarr), T.ip(opts), upscaler.encode(),
                                plane_upscaler.encode(), b"", T.fp(peak_out))
        assert r == 0, f"pl_render_image failed / degraded: {r}"
        if tgt.planes:
            return [gpu.download(t, fmt, *tgt.plane_size(i))
                    for i, (t, (fmt, _, _)) in enumerate(zip(ptex, tgt.planes))], peak_out
        return gpu.download(dst, tgt.fmt, tgt.w, tgt.h), peak_out
    finally:
        gpu.free(dst, *texs, *ptex)


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
    # stated tolerance: +-1 ULP of the 16-bit output; achieved: bit-exact (same fp32 operation
    # order, same deterministic coordinates)
    assert ulp.max() == 0, (ulp.max(), np.mean(ulp > 0))


def test_c2_with_deband(gpu):
    gpu.lib.shim_renderer_reset(C.c_void_p(gpu.ctx))
    planes = T.yuv420p10(192, 108, seed=4)
    img = P.Frame(planes, T.SYS["BT_709"], csp=SDR)
    tgt = P.Target(384, 216, "rgba16hf", csp=SDR)
    got, _ = gpu_render(gpu, img, tgt, upscaler="ewa_lanczos", deband=True, frame_index=5)
    want, _ = P.render(img, tgt, upscaler="ewa_lanczos", deband=True, frame_index=5)
    ulp = T.ulp_diff_half(got, want)
    assert ulp.max() == 0, (ulp.max(), np.mean(ulp > 0))


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
    # peak statistics are integer sums over work groups: the smoothed state must be identical
    assert np.float32(peak[0]) == np.float32(pstate[0]) and np.float32(peak[1]) == np.float32(pstate[1]), (peak, pstate)
    d = np.abs(got[..., :3].astype(np.int32) - want[..., :3].astype(np.int32))
    # stated tolerance +-1 LSB of the 16-bit output; achieved: identical integers
    assert d.max() == 0, (d.max(), np.mean(d > 0))


def test_c3_no_deband_noise_content(gpu):
    """Saturated random content (worst case for the PQ / gamut-mapping arithmetic)."""
    gpu.lib.shim_renderer_reset(C.c_void_p(gpu.ctx))
    planes = T.yuv420p10(256, 144, seed=8, kind="noise")
    img = P.Frame(planes, T.SYS["BT_2020_NC"], csp=HDR10)
    tgt = P.Target(256, 144, "rgba16", sample_depth=16, color_depth=16, csp=SDR)
    kw = dict(upscaler="ewa_lanczos", peak=2, cmap_hq=True)
    got, peak = gpu_render(gpu, img, tgt, **kw)
    want, pstate = P.render(img, tgt, **kw)
    d = np.abs(got[..., :3].astype(np.int32) - want[..., :3].astype(np.int32))
    assert d.max() == 0, (d.max(), np.mean(d > 0))


def test_plain_render_matches(gpu):
    planes = T.yuv420p10(64, 36, seed=1)
    img = P.Frame(planes, T.SYS["BT_709"], csp=SDR)
    tgt = P.Target(64, 36, "rgba16hf", csp=SDR)
    got, _ = gpu_render(gpu, img, tgt)
    want, _ = P.render(img, tgt)
    assert T.ulp_diff_half(got, want).max() == 0


@pytest.mark.parametrize("size", [(192, 108, 384, 216), (200, 120, 400, 240), (160, 90, 250, 141)])
def test_hdr_upscale_with_peak(gpu, size):
    """HDR10 source that is also upscaled: the peak is measured on the source-resolution merged
    pass (renderer.c:1460-1470), the FBO round trip feeds the main EWA scaler, then the colour
    mapping runs with the same-frame peak. Odd sizes exercise the generic polar kernel and
    partially filled work groups."""
    w, h, ow, oh = size
    gpu.lib.shim_renderer_reset(C.c_void_p(gpu.ctx))
    planes = T.yuv420p10(w, h, seed=11)
    img = P.Frame(planes, T.SYS["BT_2020_NC"], csp=HDR10)
    tgt = P.Target(ow, oh, "rgba16", sample_depth=16, color_depth=16, csp=SDR)
    kw = dict(upscaler="ewa_lanczos", peak=2, cmap_hq=True)
    got, peak = gpu_render(gpu, img, tgt, **kw)
    want, pstate = P.render(img, tgt, **kw)
    assert np.float32(peak[0]) == np.float32(pstate[0]) and np.float32(peak[1]) == np.float32(pstate[1])
    d = np.abs(got[..., :3].astype(np.int32) - want[..., :3].astype(np.int32))
    exact_coords = (ow, oh) == (2 * w, 2 * h)
    if exact_coords:
        assert d.max() == 0, (d.max(), np.mean(d > 0))
    else:
        # generic ratios: sample positions are fp32 expressions evaluated in a different (but
        # equally valid) order than the oracle's exact evaluation -> 1 LSB flips are possible
        assert d.max() <= 1 and np.mean(d > 0) < 1e-2, (d.max(), np.mean(d > 0))


def test_peak_smoothing_over_frames(gpu):
    """Three consecutive frames: the IIR-smoothed peak state (colorspace.c:1071-1180) carried by
    the renderer must track the oracle's state machine exactly, and so must every frame."""
    gpu.lib.shim_renderer_reset(C.c_void_p(gpu.ctx))
    tgt = P.Target(128, 72, "rgba16", sample_depth=16, color_depth=16, csp=SDR)
    state = None
    for i, (seed, kind) in enumerate([(20, "rings"), (21, "noise"), (22, "rings")]):
        planes = T.yuv420p10(128, 72, seed=seed, kind=kind)
        img = P.Frame(planes, T.SYS["BT_2020_NC"], csp=HDR10)
        kw = dict(peak=2, cmap_hq=True, frame_index=i)
        got, peak = gpu_render(gpu, img, tgt, **kw)
        want, state = P.render(img, tgt, peak_state=state, **kw)
        assert np.float32(peak[0]) == np.float32(state[0]) and np.float32(peak[1]) == np.float32(state[1]), (i, peak, state)
        assert np.array_equal(got[..., :3], want[..., :3]), i


@pytest.mark.parametrize("kernel", ["sierra-lite", "floyd-steinberg"])
def test_render_with_error_diffusion(gpu, kernel):
    """Config C5's output stage at reduced size: error diffusion to 8 bits replaces the dither
    matrix (renderer.c:2886-2899); two extra rgba16hf FBOs, one single-work-group dispatch."""
    planes = T.yuv420p10(160, 90, seed=14)
    img = P.Frame(planes, T.SYS["BT_709"], csp=SDR)
    tgt = P.Target(160, 90, "rgba8", sample_depth=8, color_depth=8, csp=SDR)
    got, _ = gpu_render(gpu, img, tgt, dither=1, error_diffusion=kernel)
    want, _ = P.render(img, tgt, dither=1, error_diffusion=kernel)
    assert np.array_equal(got[..., :3], want[..., :3]), np.mean(got[..., :3] != want[..., :3])


@pytest.mark.parametrize("variant", [0, 1])
def test_c5_dolby_vision_gamut_error_diffusion_pipeline(gpu, variant):
    """Config C5 at reduced size, as ONE pipeline: Dolby Vision profile-5-style source (reshaping
    curves + RPU matrices of the reference's dovi_meta fixture), EWA chroma upscale, perceptual
    gamut mapping / tone mapping to BT.709, error diffusion (sierra-lite) to 8 bits."""
    gpu.lib.shim_renderer_reset(C.c_void_p(gpu.ctx))
    planes = T.yuv420p10(320, 180, seed=31 + variant)
    img = P.Frame(planes, T.SYS["DOLBYVISION"], levels=T.LEVELS["FULL"], csp=HDR10, dovi=variant)
    tgt = P.Target(320, 180, "rgba8", sample_depth=8, color_depth=8, csp=SDR)
    kw = dict(upscaler="ewa_lanczos", dither=1, cmap_hq=True, error_diffusion="sierra-lite")
    got, _ = gpu_render(gpu, img, tgt, **kw)
    want, _ = P.render(img, tgt, **kw)
    d = np.abs(got[..., :3].astype(np.int32) - want[..., :3].astype(np.int32))
    # stated tolerance: +-1 LSB of the output; error diffusion propagates any difference, so the
    # achieved result must be (and is) identical
    assert d.max() == 0, (int(d.max()), float(np.mean(d > 0)))


@pytest.mark.parametrize("peak", [0, 2])
def test_c3_with_contrast_recovery(gpu, peak):
    """pl_render_high_quality_params' colour mapping as the reference defines it: contrast
    recovery 0.3 with the low-pass feature map (renderer.c:2089-2154 — feature extraction,
    widened bicubic down-scaling by 3.5 with mirrored borders — and the bicubic detail fetch of
    colorspace.c:1879-1921)."""
    gpu.lib.shim_renderer_reset(C.c_void_p(gpu.ctx))
    planes = T.yuv420p10(320, 180, seed=41)
    img = P.Frame(planes, T.SYS["BT_2020_NC"], csp=HDR10)
    tgt = P.Target(320, 180, "rgba16", sample_depth=16, color_depth=16, csp=SDR)
    kw = dict(upscaler="ewa_lanczos", deband=True, peak=peak, cmap_hq=True, frame_index=0,
              contrast_recovery=0.3)
    got, _ = gpu_render(gpu, img, tgt, **kw)
    want, _ = P.render(img, tgt, **kw)
    without, _ = P.render(img, tgt, **dict(kw, contrast_recovery=0.0))
    assert np.mean(want[..., :3] != without[..., :3]) > 0.2  # the feature really changes the picture
    d = np.abs(got[..., :3].astype(np.int32) - want[..., :3].astype(np.int32))
    # the separable passes evaluate their sample positions in a different (equally valid) fp32
    # order than the oracle (1 ULP), which the fp16 feature map mostly absorbs: stated tolerance
    # +-1 LSB of the 16-bit output, rare pixels
    assert d.max() <= 1 and np.mean(d > 0) < 1e-3, (int(d.max()), float(np.mean(d > 0)))


YUV420P16 = [("r16", (0,), (1, 1)), ("r16", (1,), (2, 2)), ("r16", (2,), (2, 2))]
P016 = [("r16", (0,), (1, 1)), ("rg16", (1, 2), (2, 2))]


@pytest.mark.parametrize("layout,loc,depth", [(YUV420P16, "left", 10), (YUV420P16, "center", 16),
                                              (P016, "top_left", 10), (P016, "left", 16)])
def test_planar_yuv_target(gpu, layout, loc, depth):
    """Planar render targets (pass_output_target with several planes, renderer.c:2796-2961):
    HDR10 4:2:0 in -> tone / gamut mapped BT.709 limited-range YCbCr 4:2:0 out, as three planes
    (yuv420p) or two (NV12 / P010 layout), chroma sited left / centre / top-left, dithered to 10
    bits or full 16-bit. Every plane must equal the oracle's (bit-exact)."""
    gpu.lib.shim_renderer_reset(C.c_void_p(gpu.ctx))
    planes = T.yuv420p10(160, 90, seed=21, kind="noise")
    img = P.Frame(planes, T.SYS["BT_2020_NC"], csp=HDR10)
    tgt = P.Target(160, 90, "rgba16", sys=T.SYS["BT_709"], levels=T.LEVELS["LIMITED"],
                   sample_depth=16, color_depth=depth, csp=SDR, planes=layout, chroma_loc=loc)
    kw = dict(plane_upscaler="ewa_lanczos", deband=True, peak=2, dither=1, cmap_hq=True, frame_index=2)
    got, _ = gpu_render(gpu, img, tgt, **kw)
    want, _ = P.render(img, tgt, **kw)
    assert len(got) == len(want) == len(layout)
    for i, (g, w) in enumerate(zip(got, want)):
        assert g.shape == w.shape, (i, g.shape, w.shape)
        d = np.abs(g.astype(np.int64) - w.astype(np.int64))
        # stated tolerance +-1 LSB of the 16-bit plane; achieved: bit-exact
        assert d.max() == 0, (i, int(d.max()), float(np.mean(d > 0)))
