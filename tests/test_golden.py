"""Golden fixtures generated from the compiled reference (tests/golden/gen_golden.py) against the
product's host code — runs anywhere (no reference, no GPU): filters, tone curves, gamut 3DLUT,
dither matrices, colour matrices and the AV1 grain tables must be reproduced bit for bit."""
import ctypes as C
import os

import numpy as np
import pytest

import plbtest as T

G = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return np.load(os.path.join(G, name))


def test_filters(pt0):
    g = load("filters.npz")
    for name, polar in (("ewa_lanczos", True), ("ewa_lanczossharp", True), ("lanczos", False),
                        ("mitchell", False), ("hermite", False), ("gaussian", False)):
        f = pt0.filter_generate(name, polar=polar)
        assert np.array_equal(f["weights"], g[name + "_weights"]), name
        geom = g[name + "_geom"]
        assert f["radius"] == geom[0] and f["radius_zero"] == geom[1], name


def test_tone_curves(pt0):
    g = load("tone.npz")
    imin, imax, iavg, omin, omax = g["range"]
    hdr = T.csp(T.PRIM["BT_2020"], T.TRC["PQ"], max_luma=1000)
    for fn in T.TONE_FNS:
        lut = pt0.tone_map_generate(fn, imin, imax, iavg, omin, omax, hdr_csp=hdr)["lut"]
        assert np.array_equal(lut, g[fn]), fn
    measured = hdr.copy()
    measured[19], measured[18] = 0.7688, 0.99999
    lut = pt0.tone_map_generate("spline", imin, np.float32(0.99999), np.float32(0.7688), omin, omax,
                                hdr_csp=measured)["lut"]
    assert np.array_equal(lut, g["spline_measured"])


def test_gamut_lut(pt0):
    g = load("gamut.npz")
    gmin, gmax = g["luma"]
    src8, dst8 = pt0.raw_primaries(T.PRIM["BT_2020"]), pt0.raw_primaries(T.PRIM["BT_709"])
    dst8c, _ = pt0.primaries_clip(dst8, src8)
    lut = pt0.gamut_map_generate("perceptual", src8, dst8c, gmin, gmax)["lut"]
    assert np.array_equal(lut.reshape(256, 32, 48, 3)[::8], g["perceptual_2020_709_h8"])


def test_dither_matrices(pt0):
    g = load("dither.npz")
    assert np.array_equal(pt0.dither_matrix(0, 64, fresh=True), g["blue_noise_64"])
    assert np.array_equal(pt0.dither_matrix(1, 64, fresh=True), g["bayer_64"])


def test_color_matrices(pt0):
    g = load("color.npz")
    for sysn, lv, sd, cd in (("BT_709", "LIMITED", 16, 10), ("BT_2020_NC", "LIMITED", 16, 10),
                              ("BT_601", "FULL", 8, 8), ("YCGCO", "FULL", 8, 8)):
        d = pt0.repr_decode(T.SYS[sysn], T.LEVELS[lv], sd, cd)
        want = g[f"decode_{sysn}_{lv}_{sd}_{cd}"]
        assert np.array_equal(np.concatenate([d["m"], d["c"], [d["scale"]]]).astype(np.float32), want)
    for p in ("BT_709", "BT_2020", "DISPLAY_P3"):
        raw = pt0.raw_primaries(T.PRIM[p])
        assert np.array_equal(raw, g["prim_" + p])
        for which in range(4):
            assert np.array_equal(pt0.matrix(which, raw), g[f"matrix{which}_{p}"]), (p, which)


def test_av1_grain_tables():
    from test_grain_tables import product_tables
    g = load("av1_grain.npz")
    gy, _, _, off, sc = product_tables(1234, 10, 0, 256, 144)
    assert np.array_equal(gy, g["plane0_desc0"])
    assert np.array_equal(off.ravel(), g["plane0_desc1"])
    assert np.array_equal(sc[0], g["plane0_desc2"])
    _, gcb, gcr, off_c, sc_c = product_tables(1234, 10, 1, 128, 72)
    uv = g["plane1_desc0"].reshape(-1, 2)
    assert np.array_equal(gcb, uv[:, 0]) and np.array_equal(gcr, uv[:, 1])
    assert np.array_equal(off_c.ravel(), g["plane1_desc1"])
    assert np.array_equal(sc_c[1], g["plane1_desc2"]) and np.array_equal(sc_c[2], g["plane1_desc3"])


def test_h274_grain_database():
    """The 169 band-limited noise blocks of SMPTE RDD 5 as the reference's sh_lut fill callback
    produced them (film_grain_h274.c:33-125); golden stores value * 255 as int8."""
    from test_grain_tables import product_h274_db
    want = (load("h274_db.npz")["db_times_255"].astype(np.float64) / 255.0).astype(np.float32)
    assert np.array_equal(product_h274_db().view(np.uint32), want.view(np.uint32))
