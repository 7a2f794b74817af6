"""The reference's known-answer tests for the pixel path (see tests/test_oracle_kat.py for the
citations) executed through the C ABI on the GPU, with the reference's own tolerances."""
import ctypes as C

import numpy as np
import pytest

import plbtest as T
from test_oracle_kat import FBO_H, FBO_W, pattern

pytestmark = pytest.mark.gpu


def run(gpu, src, ops, oargs, want_peak=False):
    h, w = src.shape[:2]
    return gpu.simple("direct", src, "rgba32f", w, h, ops=ops, oargs=oargs, want_peak=want_peak)


def test_peak_detection_gradient(gpu):
    gpu.lib.shim_renderer_reset(C.c_void_p(gpu.ctx))
    src = pattern()
    t0 = T.t0()
    _, peak = run(gpu, src, [9], [T.csp(0, T.TRC["GAMMA22"]), [0.0]], want_peak=True)
    lin = (0.212639 * src[..., 0].astype(np.float64) ** 2.2 + 0.715169 * src[..., 1].astype(np.float64) ** 2.2)
    pq = np.array([t0.hdr_rescale(T.HDR["NORM"], T.HDR["PQ"], float(v)) for v in lin.ravel()])
    assert abs(peak[1] - pq.max()) < 1e-4 and abs(peak[0] - pq.mean()) < 1e-3


@pytest.mark.parametrize("trc", [k for k in T.TRC if k != "UNKNOWN"])
def test_transfer_round_trip(gpu, trc):
    csp = T.csp(0, T.TRC[trc], min_luma=1e-6)
    src = pattern()
    got = run(gpu, src, [4, 3], [csp, csp])
    hdr = trc in ("PQ", "HLG", "V_LOG", "S_LOG1", "S_LOG2", "SCRGB")
    assert np.abs(got - src).max() < (1e-4 if hdr else 1e-6), trc


@pytest.mark.parametrize("gamut,ref", [
    ("clip", [[2.5375, 2.5375, 2.5375], [4.2135, -0.3160, -0.0461], [-1.4911, 2.8747, -0.2552],
              [-0.1849, -0.0212, 2.8388]]),
    ("saturation", [[2.5375, 2.5375, 2.5375], [3.1083, -0.1067, -0.0464], [-0.5708, 2.6442, -0.1995],
                    [0.0, 0.0, 2.7834]]),
])
def test_scrgb_gamut_goldens(gpu, gamut, ref):
    src_csp = T.csp(T.PRIM["BT_2020"], T.TRC["LINEAR"])
    dst_csp = T.csp(T.PRIM["BT_709"], T.TRC["SCRGB"], hdr_prim=T.t0().raw_primaries(T.PRIM["DISPLAY_P3"]))
    rgb = np.array([[[1, 1, 1, 1], [1, 0, 0, 1], [0, 1, 0, 1], [0, 0, 1, 1]]], np.float32)
    oa = [src_csp, dst_csp, [0.0, T.TONE_FNS.index("clip"), T.GAMUT_FNS.index(gamut), 1.0]]
    got = run(gpu, rgb, [5], oa)
    assert np.abs(got[0, :, :3] - np.array(ref, np.float32)).max() < 1e-4


@pytest.mark.parametrize("sys", [k for k in T.SYS if k not in ("DOLBYVISION", "BT_2100_PQ", "BT_2100_HLG")])
def test_color_system_round_trip(gpu, sys):
    src = pattern()
    r = [T.SYS[sys], 0, 0, 0, 0]
    got = run(gpu, src, [2, 1], [r, r])
    eps = 1e-5 if sys in ("BT_2020_C", "XYZ") else 1e-6
    assert np.abs(got - src).max() < eps, sys
