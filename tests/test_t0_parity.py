"""Host-side parameter derivation of libplacebo_b200 (csrc/t0/*.c) against the UNMODIFIED
reference built into oracle/_ref (skipped where the reference tree was never available; the
committed goldens in tests/golden cover that case, see test_golden.py). Everything is compared
bit for bit."""
import ctypes as C

import numpy as np
import pytest

import plbtest as T

FILTERS = ["ewa_lanczos", "ewa_lanczossharp", "ewa_lanczos4sharpest", "ewa_jinc", "ewa_hann",
           "ewa_ginseng", "ewa_robidoux", "ewa_robidouxsharp", "lanczos", "ginseng", "sinc",
           "spline16", "spline36", "spline64", "bicubic", "hermite", "mitchell",
           "mitchell_clamp", "catmull_rom", "robidoux", "robidouxsharp", "gaussian",
           "bilinear", "nearest"]


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("name", FILTERS)
@pytest.mark.parametrize("blur", [0.0, 0.9, 2.0])
def test_filter_luts(ref, pt0, name, blur):
    if name == "nearest" and 0 < blur < 1:
        pytest.skip("the reference itself asserts (wsum > 0, filters.c:174) for this input")
    a, b = ref.filter_generate(name, blur), pt0.filter_generate(name, blur)
    assert a is not None and b is not None
    for k in ("radius", "radius_zero", "row_size", "row_stride", "polar", "antiring"):
        assert a[k] == b[k], k
    assert np.array_equal(bits(a["weights"]), bits(b["weights"]))


def test_filter_sample_points(ref, pt0):
    for name in FILTERS:
        for x in np.linspace(0, 4.5, 37):
            assert ref.filter_sample(name, x) == pt0.filter_sample(name, x), (name, x)


def test_unknown_filter_is_rejected(pt0):
    assert pt0.filter_generate("no_such_filter") is None


BITS = [(8, 8, 0), (16, 10, 0), (16, 10, 6), (16, 12, 0), (10, 10, 0), (0, 0, 0), (16, 16, 0), (16, 8, 0)]


@pytest.mark.parametrize("sys", [s for s in range(14) if s != T.SYS["DOLBYVISION"]])
def test_repr_decode_encode(ref, pt0, sys):
    for lev in (0, 1, 2):
        for b in BITS:
            assert np.array_equal(bits(ref.repr_decode(sys, lev, *b)["raw"]),
                                  bits(pt0.repr_decode(sys, lev, *b)["raw"])), (sys, lev, b)
            assert np.array_equal(bits(ref.repr_encode(sys, lev, *b)["raw"]),
                                  bits(pt0.repr_encode(sys, lev, *b)["raw"])), (sys, lev, b)


def test_primaries_and_matrices(ref, pt0):
    prims = [ref.raw_primaries(p) for p in range(18)]
    for p in range(18):
        assert np.array_equal(prims[p], pt0.raw_primaries(p))
        for which in range(6):
            assert np.array_equal(bits(ref.matrix(which, prims[p])), bits(pt0.matrix(which, prims[p])))
    for s in (3, 6, 11, 9):
        for d in (3, 6, 12):
            for intent in (-1, 0, 1, 2, 3):
                assert np.array_equal(bits(ref.color_mapping_matrix(prims[s], prims[d], intent)),
                                      bits(pt0.color_mapping_matrix(prims[s], prims[d], intent)))
            a, ca = ref.primaries_clip(prims[s], prims[d])
            b, cb = pt0.primaries_clip(prims[s], prims[d])
            assert ca == cb and np.array_equal(bits(a), bits(b))


def test_hdr_rescale(ref, pt0):
    xs = [0.0, 1e-6, 1e-3, 0.01, 0.1, 0.5, 0.58, 0.75, 1.0, 10.0, 100.0, 203.0, 1000.0, 10000.0]
    for a in range(4):
        for b in range(4):
            for x in xs:
                r, p = ref.hdr_rescale(a, b, x), pt0.hdr_rescale(a, b, x)
                assert r == p or (np.isnan(r) and np.isnan(p)), (a, b, x)


CSPS = [
    T.csp(), T.csp(T.PRIM["BT_709"], T.TRC["BT_1886"]), T.csp(T.PRIM["BT_2020"], T.TRC["PQ"]),
    T.csp(T.PRIM["BT_2020"], T.TRC["PQ"], min_luma=0.005, max_luma=1000),
    T.csp(T.PRIM["BT_2020"], T.TRC["HLG"]), T.csp(T.PRIM["DISPLAY_P3"], T.TRC["SRGB"]),
    T.csp(T.PRIM["BT_2020"], T.TRC["PQ"], max_luma=4000, max_cll=1200, max_fall=300),
    T.csp(T.PRIM["BT_2020"], T.TRC["PQ"], max_luma=1000, max_pq_y=0.62, avg_pq_y=0.31),
    T.csp(T.PRIM["BT_2020"], T.TRC["PQ"], max_luma=1000, scene_max=(800, 900, 400), scene_avg=120),
    T.csp(T.PRIM["BT_709"], T.TRC["GAMMA22"], max_luma=400), T.csp(0, T.TRC["LINEAR"]),
    T.csp(T.PRIM["BT_709"], T.TRC["SCRGB"]),
]


def test_color_space_inference_and_nominal_luma(ref, pt0):
    for c in CSPS:
        assert np.array_equal(bits(ref.csp_infer(c)), bits(pt0.csp_infer(c)))
        for md in range(5):
            for sc in range(4):
                assert np.array_equal(bits(ref.nominal_luma(c, md, sc)), bits(pt0.nominal_luma(c, md, sc)))
        for d in CSPS:
            ra, rb = ref.csp_infer_map(c, d)
            pa, pb = pt0.csp_infer_map(c, d)
            assert np.array_equal(bits(ra), bits(pa)) and np.array_equal(bits(rb), bits(pb))


RANGES = [(1e-4, 0.75, 0.0, 0.02, 0.58), (1e-4, 0.75, 0.3, 0.02, 0.58), (1e-4, 0.9, 0.0, 2e-4, 0.58),
          (0.02, 0.58, 0.0, 1e-4, 0.75), (0.01, 0.6, 0.25, 0.01, 0.5), (1e-4, 0.58, 0.1, 1e-4, 0.58)]


@pytest.mark.parametrize("fn", T.TONE_FNS)
def test_tone_map_luts(ref, pt0, fn):
    for r in RANGES:
        for size in (256, 17):
            a = ref.tone_map_generate(fn, *r, lut_size=size, hdr_csp=CSPS[2])
            b = pt0.tone_map_generate(fn, *r, lut_size=size, hdr_csp=CSPS[2])
            assert a["noop"] == b["noop"]
            assert np.array_equal(bits(a["params"]), bits(b["params"]))
            assert np.array_equal(bits(a["lut"]), bits(b["lut"])), (fn, r)


@pytest.mark.parametrize("fn", T.GAMUT_FNS)
def test_gamut_map_luts(ref, pt0, fn):
    p709, p2020, p3 = (ref.raw_primaries(T.PRIM[k]) for k in ("BT_709", "BT_2020", "DISPLAY_P3"))
    for s, d in ((p2020, p709), (p709, p2020), (p3, p709), (p709, p709)):
        a = ref.gamut_map_generate(fn, s, d, 0.02, 0.58, size=(16, 12, 40))
        b = pt0.gamut_map_generate(fn, s, d, 0.02, 0.58, size=(16, 12, 40))
        assert a["noop"] == b["noop"]
        assert np.array_equal(bits(a["lut"]), bits(b["lut"])), fn


def test_default_gamut_lut_full_size(ref, pt0):
    """The 48x32x256 perceptual LUT of config C3 (BT.2020 -> BT.709)."""
    p709, p2020 = ref.raw_primaries(T.PRIM["BT_709"]), ref.raw_primaries(T.PRIM["BT_2020"])
    lo, hi = ref.hdr_rescale(T.HDR["NITS"], T.HDR["PQ"], 0.203), ref.hdr_rescale(T.HDR["NITS"], T.HDR["PQ"], 203.0)
    a = ref.gamut_map_generate("perceptual", p2020, p709, lo, hi)
    b = pt0.gamut_map_generate("perceptual", p2020, p709, lo, hi)
    assert np.array_equal(bits(a["lut"]), bits(b["lut"]))


def test_dither_matrices(pt0):
    """Bayer always; blue noise against a FRESH reference process (the reference consumes
    libc rand() state, so only its first call per process is reproducible)."""
    import subprocess, sys, os
    code = ("import sys; sys.path.insert(0, %r); import plbtest as T, numpy as np;"
            "r = T.ref(); sys.stdout.buffer.write(r.dither_matrix(0, int(sys.argv[1])).tobytes())")
    if not T.have_ref():
        pytest.skip("no reference")
    for size in (8, 64):
        assert np.array_equal(T.ref().dither_matrix(1, size), pt0.dither_matrix(1, size))
    for size in (16, 64):
        out = subprocess.run([sys.executable, "-c", code % os.path.join(T.ROOT, "tests"), str(size)],
                             capture_output=True, check=True).stdout
        fresh = np.frombuffer(out, np.float32).reshape(size, size)
        assert np.array_equal(fresh, pt0.dither_matrix(0, size))
        # and the product is deterministic across calls, which the reference is not
        assert np.array_equal(pt0.dither_matrix(0, size), pt0.dither_matrix(0, size))


def test_error_diffusion_kernels(ref, pt0):
    for name in ["simple", "false-fs", "sierra-lite", "floyd-steinberg", "atkinson",
                 "jarvis-judice-ninke", "stucki", "burkes", "sierra-2", "sierra-3"]:
        a, b = ref.error_diffusion_kernel(name), pt0.error_diffusion_kernel(name)
        assert a["shift"] == b["shift"] and a["divisor"] == b["divisor"]
        assert np.array_equal(a["pattern"], b["pattern"])
    assert pt0.error_diffusion_kernel("nope") is None


def test_dovi_repr_decode(ref, pt0):
    """pl_color_repr_normalize + pl_color_repr_decode for a Dolby Vision repr (RPU matrix and
    offsets) and the LMS->RGB x linear product: product T0 == compiled reference."""
    for variant in (0, 1):
        m = T.dovi_fixture(variant)
        args = (16, 10, list(m.nonlinear_offset), list(m.nonlinear), list(m.linear))
        a, b = ref.repr_decode_dovi(*args), pt0.repr_decode_dovi(*args)
        for k in a:
            assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), (variant, k)


# ---- colorspace.h surface beyond the render path ---------------------------------------------

def _both(ref, pt0, name, restype=None):
    fr, fp_ = ref.f(name), pt0.f(name)
    if restype is not None:
        fr.restype = fp_.restype = restype
    return fr, fp_





def test_filter_function_table(ref, pt0):
    """Every kernel / window the product exports exists under the same name in the reference
    (the reference additionally keeps deprecated aliases of the cubic family)."""
    cr, cp = _both(ref, pt0, "num_filter_functions")
    nr, np_ = _both(ref, pt0, "filter_function_name", C.c_char_p)
    ref_names = {nr(i) for i in range(cr())}
    ours = [np_(i) for i in range(cp())]
    assert len(ours) == len(set(ours)) >= 20
    assert set(ours) <= ref_names, set(ours) - ref_names



def test_filter_presets(ref, pt0):
    """pl_filter_presets: same names, order, configurations and descriptions as the reference;
    pl_find_filter_preset finds each of them."""
    fr, fp_ = _both(ref, pt0, "filter_preset")
    i = 0
    while True:
        a, b = C.create_string_buffer(256), C.create_string_buffer(256)
        ra, rb = fr(i, a, 256), fp_(i, b, 256)
        assert (ra < 0) == (rb < 0), i
        if ra < 0:
            break
        assert a.value == b.value and a.value.endswith(b"|1"), (i, a.value, b.value)
        i += 1
    assert i >= 26
    assert not pt0.lib.plt_filter_preset(-1, b, 256) >= 0


