"""Colour shaders (decode, linearize/delinearize, sigmoid, colour map, dither, peak detection)
through the C ABI against the CPU oracle. Input and output are rgba32f so any difference is
arithmetic, not storage rounding. Kernels and oracle share one definition of every
transcendental function, so most cases are bit-identical; the tolerance written
in each test is the stated bound (+-1 LSB of a 16-bit output), not an allowance for libm."""
import numpy as np
import pytest

import pipeline_oracle as P
import plbtest as T

pytestmark = pytest.mark.gpu

W, H = 96, 64


def rgb(seed, lo=0.0, hi=1.0, gray_frac=0.2):
    r = np.random.default_rng(seed)
    a = r.uniform(lo, hi, (H, W, 4)).astype(np.float32)
    g = r.random((H, W)) < gray_frac
    a[g, 1] = a[g, 0]
    a[g, 2] = a[g, 0]
    a[..., 3] = 1.0
    a[0, :8, :3] = 0.0
    a[0, 8:16, :3] = hi
    return a


def run_ops(gpu, src, ops, oargs, want_peak=False):
    return gpu.simple("direct", src, "rgba32f", W, H, ops=ops, oargs=oargs, want_peak=want_peak)


def relerr(got, want, floor=1e-4):
    got, want = got[..., :3].astype(np.float64), want[..., :3].astype(np.float64)
    return np.abs(got - want) / np.maximum(np.abs(want), floor)


@pytest.mark.parametrize("trc", [k for k in T.TRC if k != "UNKNOWN"])
def test_linearize_delinearize(gpu, oracle, trc):
    t = T.TRC[trc]
    csp = T.csp(T.PRIM["BT_709"], t)
    c2 = T.t0().csp_infer(csp)
    smin, smax, _ = T.t0().nominal_luma(c2, T.META["HDR10"], T.HDR["NORM"])
    luma = P.luma_coeffs(T.t0(), int(c2[0]))
    src = rgb(1)
    got = run_ops(gpu, src, [3], [csp])
    want = oracle.linearize(src, t, smin, smax, luma)
    e = relerr(got, want, 1e-5)
    assert e.max() < 2e-5, (trc, e.max())
    lin = want.copy()
    got2 = run_ops(gpu, lin, [4], [csp])
    # pl_color_space_is_black_scaled (colorspace.c:525-554) minus HLG (shaders/colorspace.c:740)
    black_scaled = trc not in ("BT_1886", "PQ", "SCRGB", "V_LOG", "S_LOG1", "S_LOG2", "HLG")
    want2 = oracle.delinearize(lin, t, smin, smax, luma, black_scaled=black_scaled)
    e2 = relerr(got2, want2, 1e-4)
    assert e2.max() < 5e-5, (trc, e2.max())


@pytest.mark.parametrize("sys,levels,depth", [("BT_709", "LIMITED", 10), ("BT_2020_NC", "LIMITED", 10),
                                              ("BT_601", "FULL", 8), ("YCGCO", "FULL", 8),
                                              ("RGB", "LIMITED", 8), ("BT_2100_PQ", "LIMITED", 10),
                                              ("BT_2100_HLG", "LIMITED", 10), ("BT_2020_C", "LIMITED", 10),
                                              ("XYZ", "FULL", 12)])
def test_decode_encode_color(gpu, oracle, sys, levels, depth):
    s, l = T.SYS[sys], T.LEVELS[levels]
    src = rgb(2, gray_frac=0.0)
    got = run_ops(gpu, src, [1], [[s, l, 16, depth, 0]])
    want = P.decode_color(oracle, T.t0(), src, s, l, (16, depth, 0))
    ok = np.isfinite(want[..., :3])
    assert np.array_equal(ok, np.isfinite(got[..., :3]))
    assert relerr(got, want, 1.0)[ok].max() < 1e-5, sys
    if sys == "XYZ":
        return
    lin = np.clip(np.nan_to_num(want), 0, 1)
    got2 = run_ops(gpu, lin, [2], [[s, l, 16, depth, 0]])
    want2 = P.encode_color(oracle, T.t0(), lin, s, l, (16, depth, 0))
    assert relerr(got2, want2, 1.0).max() < 1e-5, sys


def test_sigmoid(gpu, oracle):
    src = rgb(3)
    got = run_ops(gpu, src, [7], [[0.75, 6.5]])
    want = oracle.sigmoidize(src)
    assert relerr(got, want).max() < 1e-5
    got = run_ops(gpu, want, [8], [[0.75, 6.5]])
    want2 = oracle.sigmoidize(want, inverse=True)
    assert relerr(got, want2).max() < 1e-5


HDR10 = T.csp(T.PRIM["BT_2020"], T.TRC["PQ"], max_luma=1000)
SDR = T.csp(T.PRIM["BT_709"], T.TRC["BT_1886"])
P3PQ = T.csp(T.PRIM["DISPLAY_P3"], T.TRC["PQ"], max_luma=600) if "DISPLAY_P3" in T.PRIM else None


def _cmap(gpu, oracle, src_csp, dst_csp, hq, tone, gamut, src, state=True):
    t0 = T.t0()
    oa = [src_csp, dst_csp, [1.0 if hq else 0.0,
                              T.TONE_FNS.index(tone) if tone else -1,
                              T.GAMUT_FNS.index(gamut) if gamut else -1, 1.0 if state else 0.0]]
    got = run_ops(gpu, src, [5], oa)
    cm = P.color_map_params(t0, src_csp.copy(), dst_csp.copy(), hq=hq, tone_fn=tone, gamut_fn=gamut,
                            state=state)
    want = oracle.color_map(src, cm["params"]) if cm is not None else src
    return got, want


@pytest.mark.parametrize("hq", [False, True])
def test_color_map_hdr10_to_sdr_default(gpu, oracle, hq):
    # linear light relative to SDR white; HDR10 1000 nits -> up to ~4.9
    src = rgb(4)  # PQ-encoded signal
    got, want = _cmap(gpu, oracle, HDR10, SDR, hq, None, None, src)
    d = np.abs(got[..., :3] - want[..., :3])
    assert d.max() < 2e-4, (d.max(), np.mean(d > 1e-5))


@pytest.mark.parametrize("tone", T.TONE_FNS)
def test_color_map_tone_functions(gpu, oracle, tone):
    src = rgb(5)
    got, want = _cmap(gpu, oracle, HDR10, SDR, False, tone, "clip", src)
    d = np.abs(got[..., :3] - want[..., :3])
    assert d.max() < 2e-4, (tone, d.max())


@pytest.mark.parametrize("gamut", T.GAMUT_FNS)
def test_color_map_gamut_functions(gpu, oracle, gamut):
    src = rgb(6, 0.0, 1.0)
    s = T.csp(T.PRIM["BT_2020"], T.TRC["BT_1886"])
    got, want = _cmap(gpu, oracle, s, SDR, False, "clip", gamut, src)
    d = np.abs(got[..., :3] - want[..., :3])
    assert d.max() < 2e-4, (gamut, d.max())


def test_color_map_without_state_object(gpu, oracle):
    """No state object: LUT methods are replaced by linear / saturation (colorspace.c:1746)."""
    src = rgb(10)
    got, want = _cmap(gpu, oracle, HDR10, SDR, False, "spline", "perceptual", src, state=False)
    d = np.abs(got[..., :3] - want[..., :3])
    assert d.max() < 2e-4, d.max()


def test_color_map_inverse_and_noop(gpu, oracle):
    src = rgb(7)
    got, want = _cmap(gpu, oracle, SDR, SDR, False, None, None, src)
    assert np.array_equal(got, want)
    got, want = _cmap(gpu, oracle, SDR, HDR10, False, None, None, src)
    d = np.abs(got[..., :3] - want[..., :3])
    assert d.max() < 2e-4, d.max()


@pytest.mark.parametrize("method,depth", [(0, 8), (1, 8), (2, 8), (3, 8), (0, 10), (0, 6)])
def test_dither(gpu, oracle, method, depth):
    """Methods: 0 blue noise, 1 ordered LUT, 2 ordered fixed, 3 white noise."""
    gpu.set_frame_index(0)
    src = rgb(8)
    got = run_ops(gpu, src, [6], [[depth, method, 6, 0, 0]])
    lut = None
    if method in (0, 1):
        lut = T.t0().dither_matrix(method, 64)
    want = oracle.dither(src, depth, method=method, lut=lut)
    scale = float((1 << depth) - 1)
    gq, wq = np.round(got[..., :3] * scale), np.round(want[..., :3] * scale)
    assert np.mean(gq != wq) < 1e-4, (method, depth, np.mean(gq != wq))
    assert np.abs(gq - wq).max() <= 1


@pytest.mark.parametrize("hq", [False, True])
def test_detect_peak(gpu, oracle, hq):
    import ctypes as C
    gpu.lib.shim_renderer_reset(C.c_void_p(gpu.ctx))
    t0 = T.t0()
    csp = t0.csp_infer(HDR10)
    smin, smax, _ = t0.nominal_luma(csp, T.META["HDR10"], T.HDR["NORM"])
    luma = P.luma_coeffs(t0, int(csp[0]))
    src = rgb(9)  # PQ-encoded
    _, peak = run_ops(gpu, src, [9], [HDR10, [1.0 if hq else 0.0]], want_peak=True)
    buf = oracle.detect_peak(src, T.TRC["PQ"], smin, smax, luma, 1.0, hq, (16, 16))
    st, used = oracle.peak_update((0.0, 0.0), buf, percentile=99.995 if hq else 100.0)
    assert abs(peak[0] - st[0]) < 2e-5 and abs(peak[1] - st[1]) < 2e-5, (peak, st)


@pytest.mark.parametrize("fn", range(6))
def test_deterministic_math_matches_oracle_definition(gpu, oracle, fn):
    """The kernels' pow/exp/log/sin/cos/atan2 (csrc/plb_math.cuh) and the oracle's
    (oracle/pl_oracle_math.h) are the same binary64 algorithm: results must be bit-identical."""
    import ctypes as C
    r = np.random.default_rng(fn)
    n = 1 << 20
    if fn == 0:
        a = np.exp(r.uniform(-30, 30, n)).astype(np.float32)
        b = r.uniform(-8, 8, n).astype(np.float32)
        sp = np.array([0, 1, -1, np.inf, -np.inf, np.nan, 0.5, 2, 1e-40, 3e38], np.float32)
        A, B = np.meshgrid(sp, sp)
        a[:100], b[:100] = A.ravel(), B.ravel()
    elif fn == 1:
        a = r.uniform(-100, 100, n).astype(np.float32); b = a
    elif fn == 2:
        a = np.exp(r.uniform(-80, 80, n)).astype(np.float32); b = a
        a[:4] = [0, -1, np.inf, np.nan]
    elif fn in (3, 4):
        a = r.uniform(-50, 50, n).astype(np.float32); b = a
    else:
        a = r.normal(0, 1, n).astype(np.float32); b = r.normal(0, 1, n).astype(np.float32)
        a[:4], b[:4] = [0, 0, np.inf, -0.0], [0, -1, np.inf, -0.0]
    got = np.empty(n, np.float32)
    ok = gpu.lib.shim_math_eval(C.c_void_p(gpu.ctx), fn, T.fp(a), T.fp(b), T.fp(got), C.c_size_t(n))
    assert ok == 0
    want = np.empty(n, np.float32)
    oracle.lib.orc_math(fn, T.fp(a), T.fp(b), T.fp(want), C.c_size_t(n))
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("variant", [0, 1])
def test_dovi_reshape(gpu, oracle, variant):
    """pl_shader_dovi_reshape with the reference's fixture: 8-piece polynomial luma, order-3 MMR
    chroma (config C5's curve set). Same fp32 operation order -> identical bits."""
    src = rgb(12, gray_frac=0.0)
    src[1, :16, 0] = np.linspace(-0.1, 1.1, 16)  # outside [0,1] and across every pivot
    got = run_ops(gpu, src, [10], [[float(variant)]])
    want = oracle.dovi_reshape(src, T.dovi_fixture(variant))
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


@pytest.mark.parametrize("variant", [0, 1])
def test_dovi_decode_color(gpu, oracle, variant):
    """pl_shader_decode_color for PL_COLOR_SYSTEM_DOLBYVISION (colorspace.c:285-310, 392-422):
    normalise, reshape, RPU matrix + offsets, PQ EOTF, LMS->RGB, PQ OETF."""
    src = rgb(13, gray_frac=0.0)
    src[..., :3] *= np.float32(1023.0 / 65535.0)  # 10-bit code values in 16-bit samples
    meta = T.dovi_fixture(variant)
    got = run_ops(gpu, src, [11], [[float(variant), 16.0, 10.0]])
    d = T.t0().repr_decode_dovi(16, 10, list(meta.nonlinear_offset), list(meta.nonlinear), list(meta.linear))
    x = oracle.scale(src, float(d["scale"]), with_alpha=False)
    x = oracle.dovi_reshape(x, meta)
    x = oracle.apply_matrix(x, d["m"], d["c"])
    want = oracle.dovi_lms(x, d["lms2rgb"])
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


ED_KERNELS = ["simple", "false-fs", "sierra-lite", "floyd-steinberg", "atkinson",
              "jarvis-judice-ninke", "stucki", "burkes", "sierra-2", "sierra-3"]


@pytest.mark.parametrize("name", ED_KERNELS)
def test_error_diffusion_kernels(gpu, oracle, name):
    """pl_shader_error_diffusion (all built-in kernels) as a stand-alone compute dispatch:
    integer error words + identical float steps -> identical output."""
    import ctypes as C
    k = T.t0().error_diffusion_kernel(name)
    assert k is not None, name
    h, w = 50, 70
    r = np.random.default_rng(5)
    src = r.random((h, w, 4), dtype=np.float32)
    src[:8] = src[:8] * 0.02 + 0.5  # flat-ish area: long error chains
    s = gpu.tex("rgba32f", w, h, src)
    d = gpu.tex("rgba32f", w, h)
    try:
        rc = gpu.lib.shim_error_diffusion(C.c_void_p(gpu.ctx), C.c_void_p(s), C.c_void_p(d), 8, name.encode())
        assert rc == 0
        got = gpu.download(d, "rgba32f", w, h)
    finally:
        gpu.free(s, d)
    want = oracle.error_diffusion(src, 8, k)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), np.mean(got != want)


@pytest.mark.parametrize("depth,size,kernel", [
    (8, (33, 1100), "sierra-lite"), (6, (64, 64), "sierra-lite"), (1, (40, 40), "sierra-lite"),
    (8, (50, 2161), "sierra-lite"), (8, (24, 4320), "sierra-lite"), (8, (40, 1500), "floyd-steinberg"),
    (7, (31, 1300), "jarvis-judice-ninke"), (8, (37, 1201), "atkinson")])
def test_error_diffusion_depths_and_tall_images(gpu, oracle, depth, size, kernel):
    """More rows than one thread block has threads (> 1024): the thread-block-cluster kernel
    (one row per thread, error ring split over the CTAs' shared memories, errors crossing a band
    boundary through distributed shared memory) at 4K / 8K heights, odd heights and kernels with
    three-row footprints."""
    import ctypes as C
    k = T.t0().error_diffusion_kernel(kernel)
    w, h = size
    src = np.random.default_rng(6).random((h, w, 4), dtype=np.float32)
    s = gpu.tex("rgba32f", w, h, src)
    d = gpu.tex("rgba32f", w, h)
    try:
        assert gpu.lib.shim_error_diffusion(C.c_void_p(gpu.ctx), C.c_void_p(s), C.c_void_p(d), depth, kernel.encode()) == 0
        got = gpu.download(d, "rgba32f", w, h)
    finally:
        gpu.free(s, d)
    want = oracle.error_diffusion(src, depth, k)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_extract_features(gpu, oracle):
    """pl_shader_extract_features (colorspace.c:1383-1404) on an HDR10 PQ and an SDR source."""
    src = rgb(3)
    for csp in (T.csp(T.PRIM["BT_2020"], T.TRC["PQ"], max_luma=1000), T.csp(T.PRIM["BT_709"], T.TRC["BT_1886"])):
        got = run_ops(gpu, src, [12], [csp])
        t0 = T.t0()
        c2 = t0.csp_infer(csp)
        smin, smax, _ = t0.nominal_luma(c2, T.META["HDR10"], T.HDR["NORM"])
        lin = oracle.linearize(src, int(c2[1]), smin, smax, P.luma_coeffs(t0, int(c2[0])))
        want = oracle.extract_features(lin, t0.matrix(2, t0.raw_primaries(int(c2[0]))))
        assert np.array_equal(got[..., 1:], want[..., 1:])
        e = relerr(got[..., :1], want[..., :1], 1e-5)
        assert e.max() < 2e-6, e.max()
        if int(c2[1]) == T.TRC["PQ"]:  # no host-side libm constants on this path: exact
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
