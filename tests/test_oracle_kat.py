"""Pin the CPU oracle to every known-answer test the reference's own suite holds for the pixel
path. Each test restates the reference test's setup and tolerance:
  src/tests/gpu_tests.c:753-785  peak detection of the 16x16 gradient vs a CPU mean/max
  src/tests/gpu_tests.c:561-577  delinearize -> linearize round trip for every transfer
  src/tests/gpu_tests.c:582-667  BT.2020 linear -> scRGB/Display-P3 clip and saturation goldens
  src/tests/gpu_tests.c:673-703  encode -> decode round trip for every colour system
  src/tests/tone_mapping.c:14-22 PQ constants (hdr_rescale identities)
These run on the CPU (no GPU marker); tests/test_gpu_kat.py repeats them through the C ABI."""
import numpy as np
import pytest

import pipeline_oracle as P
import plbtest as T

FBO_W = FBO_H = 16


def pattern():
    """TEST_FBO_PATTERN (gpu_tests.c:437-454): color = ((x+.5)/W, (y+.5)/H, 0, 1)"""
    a = np.zeros((FBO_H, FBO_W, 4), np.float32)
    a[..., 0] = ((np.arange(FBO_W) + 0.5) / FBO_W)[None, :]
    a[..., 1] = ((np.arange(FBO_H) + 0.5) / FBO_H)[:, None]
    a[..., 3] = 1.0
    return a


def test_peak_detection_gradient(oracle):
    t0 = T.t0()
    src = pattern()
    csp = t0.csp_infer(T.csp(0, T.TRC["GAMMA22"]))
    smin, smax, _ = t0.nominal_luma(csp, T.META["HDR10"], T.HDR["NORM"])
    luma = P.luma_coeffs(t0, int(csp[0]))
    buf = oracle.detect_peak(src, T.TRC["GAMMA22"], smin, smax, luma, 1.0, False, (16, 16))
    (avg, mx), _ = oracle.peak_update((0.0, 0.0), buf)
    lin = (0.212639 * src[..., 0].astype(np.float64) ** 2.2 + 0.715169 * src[..., 1].astype(np.float64) ** 2.2 +
           0.072192 * src[..., 2].astype(np.float64) ** 2.2)
    pq = np.array([t0.hdr_rescale(T.HDR["NORM"], T.HDR["PQ"], float(v)) for v in lin.ravel()])
    assert abs(mx - pq.max()) < 1e-4
    assert abs(avg - pq.mean()) < 1e-3


@pytest.mark.parametrize("trc", [k for k in T.TRC if k != "UNKNOWN"])
def test_transfer_round_trip(oracle, trc):
    t0 = T.t0()
    t = T.TRC[trc]
    csp = t0.csp_infer(T.csp(0, t, min_luma=1e-6))  # PL_COLOR_HDR_BLACK
    smin, smax, _ = t0.nominal_luma(csp, T.META["HDR10"], T.HDR["NORM"])
    luma = P.luma_coeffs(t0, int(csp[0]))
    src = pattern()
    black_scaled = trc not in ("BT_1886", "PQ", "SCRGB", "V_LOG", "S_LOG1", "S_LOG2", "HLG")
    enc = oracle.delinearize(src, t, smin, smax, luma, black_scaled=black_scaled)
    dec = oracle.linearize(enc, t, smin, smax, luma)
    hdr = trc in ("PQ", "HLG", "V_LOG", "S_LOG1", "S_LOG2", "SCRGB")
    assert np.abs(dec - src).max() < (1e-4 if hdr else 1e-6), trc


@pytest.mark.parametrize("gamut,ref", [
    ("clip", [[2.5375, 2.5375, 2.5375], [4.2135, -0.3160, -0.0461], [-1.4911, 2.8747, -0.2552],
              [-0.1849, -0.0212, 2.8388]]),
    ("saturation", [[2.5375, 2.5375, 2.5375], [3.1083, -0.1067, -0.0464], [-0.5708, 2.6442, -0.1995],
                    [0.0, 0.0, 2.7834]]),
])
def test_scrgb_gamut_goldens(oracle, gamut, ref):
    t0 = T.t0()
    src = T.csp(T.PRIM["BT_2020"], T.TRC["LINEAR"])
    dst = T.csp(T.PRIM["BT_709"], T.TRC["SCRGB"], hdr_prim=t0.raw_primaries(T.PRIM["DISPLAY_P3"]))
    rgb = np.array([[[1, 1, 1, 1], [1, 0, 0, 1], [0, 1, 0, 1], [0, 0, 1, 1]]], np.float32)
    cm = P.color_map_params(t0, src, dst, tone_fn="clip", gamut_fn=gamut)
    out = oracle.color_map(rgb, cm["params"])
    assert np.abs(out[0, :, :3] - np.array(ref, np.float32)).max() < 1e-4


@pytest.mark.parametrize("sys", [k for k in T.SYS if k not in ("DOLBYVISION", "BT_2100_PQ", "BT_2100_HLG")])
def test_color_system_round_trip(oracle, sys):
    t0 = T.t0()
    s = T.SYS[sys]
    src = pattern()
    enc = P.encode_color(oracle, t0, src, s, 0, (0, 0, 0))
    dec = P.decode_color(oracle, t0, enc, s, 0, (0, 0, 0))
    eps = 1e-5 if sys in ("BT_2020_C", "XYZ") else 1e-6
    assert np.abs(dec - src).max() < eps, sys


def test_pq_constants():
    """tone_mapping.c:14-22: PQ <-> nits identities of pl_hdr_rescale"""
    t0 = T.t0()
    feq = lambda a, b, eps: abs(np.float32(a) - np.float32(b)) <= eps * max(1.0, abs(a))  # REQUIRE_FEQ, utils.h:139  # noqa: E731
    for pt in (T.t0(), T.product_t0()):
        assert feq(pt.hdr_rescale(T.HDR["PQ"], T.HDR["NITS"], 0.0), 0.0, 1e-2)
        assert feq(pt.hdr_rescale(T.HDR["PQ"], T.HDR["NITS"], 1.0), 10000.0, 1e-2)
        assert feq(pt.hdr_rescale(T.HDR["PQ"], T.HDR["NITS"], 0.58), 203.0, 1e-2)
        x = np.float32(0.0)
        while x < 1.0:  # `for (float x = 0.0f; x < 1.0f; x += 0.01f)`
            x_next = np.float32(x + np.float32(0.01))
            y = pt.hdr_rescale(T.HDR["NORM"], T.HDR["PQ"], pt.hdr_rescale(T.HDR["PQ"], T.HDR["NORM"], float(x)))
            assert feq(float(x), y, 1e-5), (x, y)
            x = x_next


@pytest.mark.parametrize("kind", ["bicubic", "hermite", "gaussian"])
def test_fast_scalers_equal_their_filter_kernels(oracle, kind):
    """The built-in fast scalers (sampling.c:319-440) fold the four taps of their kernel into
    bilinear fetches; at an exact 2x ratio their result must equal the separable convolution with
    the same pl_filter_config (cubic B-spline, Hermite, Gaussian) up to LUT / rounding noise.
    Pins the oracle's restatement of the GLSL to the reference's filter definitions."""
    rng = np.random.default_rng(0)
    src = rng.random((24, 32, 4), dtype=np.float32)
    fast = oracle.sample_fast(src, 64, 48, kind, coord_mode=T.COORD_EXACT)
    f = T.t0().filter_generate(kind, polar=False)
    v = oracle.sample_ortho(src, 32, 48, f, 0, comp_mask=15, coord_mode=T.COORD_EXACT, ratio=2.0)
    sep = oracle.sample_ortho(v, 64, 48, f, 1, comp_mask=15, coord_mode=T.COORD_EXACT, ratio=2.0)
    assert np.abs(fast - sep)[4:-4, 4:-4].max() < 2e-5


def test_oversample_is_nearest_at_integer_ratios(oracle):
    rng = np.random.default_rng(1)
    src = rng.random((24, 32, 4), dtype=np.float32)
    ov = oracle.sample_fast(src, 64, 48, "oversample", coord_mode=T.COORD_EXACT)
    nn = oracle.sample_direct(src, 64, 48, linear=False, coord_mode=T.COORD_EXACT)
    assert np.abs(ov - nn).max() < 2e-6
