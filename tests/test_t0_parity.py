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


@pytest.mark.parametrize("direction", [0, 1])
def test_cpu_transfer_functions(ref, pt0, direction):
    """pl_color_linearize / pl_color_delinearize: bit-identical with the compiled reference for
    every transfer function, SDR and HDR display ranges, values below 0 and above 1."""
    fr, fp_ = _both(ref, pt0, "color_transfer_cpu")
    rng = np.random.default_rng(5 + direction)
    base = np.concatenate([rng.random(3 * 400), rng.uniform(-0.5, 4.0, 3 * 200),
                           [0, 0, 0, 1, 1, 1, 0.5, 0.04045, 0.0031308, 1e-7, 0.181, 0.030001]]).astype(T.f32)
    for trc in range(len(T.TRC)):
        for prim in (T.PRIM["BT_709"], T.PRIM["BT_2020"]):
            for lo, hi in ((0.0, 0.0), (0.005, 600.0), (0.2, 100.0)):
                a, b = base.copy(), base.copy()
                fr(prim, trc, C.c_float(lo), C.c_float(hi), direction, T.fp(a), a.size // 3)
                fp_(prim, trc, C.c_float(lo), C.c_float(hi), direction, T.fp(b), b.size // 3)
                assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (trc, prim, lo, hi)


def test_cone_matrices_and_vision_presets(ref, pt0):
    fr, fp_ = _both(ref, pt0, "cone_matrix")
    for cones in range(8):
        for strength in (0.0, 0.25, 0.5, 1.0, 1.5):
            for prim in (T.PRIM["BT_709"], T.PRIM["BT_2020"], T.PRIM["DCI_P3"]):
                a, b = np.zeros(9, T.f32), np.zeros(9, T.f32)
                fr(cones, C.c_float(strength), prim, T.fp(a))
                fp_(cones, C.c_float(strength), prim, T.fp(b))
                assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (cones, strength, prim)
    vr, vp = _both(ref, pt0, "vision_presets")
    a, b = np.zeros(18, T.f32), np.zeros(18, T.f32)
    vr(T.fp(a))
    vp(T.fp(b))
    assert np.array_equal(a, b)


def test_white_point_guesses_names_and_repr_merge(ref, pt0):
    wr, wp = _both(ref, pt0, "white_from_temp")
    for temp in (1000, 1667, 2500, 3200, 4000, 5503, 6504, 9300, 25000, 40000):
        a, b = np.zeros(2, T.f32), np.zeros(2, T.f32)
        wr(C.c_float(temp), T.fp(a))
        wp(C.c_float(temp), T.fp(b))
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), temp
    gr, gp = _both(ref, pt0, "guess_from_size")
    for w, h in ((720, 480), (720, 486), (720, 576), (1024, 576), (1280, 720), (1920, 1080),
                 (2000, 576), (200, 200), (1279, 577), (640, 360)):
        a, b = np.zeros(2, np.int32), np.zeros(2, np.int32)
        gr(w, h, T.ip(a))
        gp(w, h, T.ip(b))
        assert np.array_equal(a, b), (w, h)
    nr, np_ = _both(ref, pt0, "enum_name", C.c_char_p)
    for kind, count in ((0, len(T.SYS)), (1, len(T.PRIM)), (2, len(T.TRC))):
        for v in range(count):
            assert nr(kind, v) == np_(kind, v) and nr(kind, v), (kind, v)
    gr2, gp2 = _both(ref, pt0, "primaries_is_wide_gamut")
    for v in range(len(T.PRIM)):
        assert bool(gr2(v)) == bool(gp2(v)), v
    mr, mp = _both(ref, pt0, "repr_merge")
    rng = np.random.default_rng(11)
    for i in range(200):
        a0 = (rng.integers(0, 3, 6) * rng.integers(0, 2, 6)).astype(np.int32)
        b0 = rng.integers(0, 3, 6).astype(np.int32)
        preset = int(rng.integers(-1, 6))
        a, b = a0.copy(), a0.copy()
        mr(T.ip(a), T.ip(b0), preset)
        mp(T.ip(b), T.ip(b0), preset)
        assert np.array_equal(a, b), (a0, b0, preset)


def test_filter_function_table(ref, pt0):
    """Every kernel / window the product exports exists under the same name in the reference
    (the reference additionally keeps deprecated aliases of the cubic family)."""
    cr, cp = _both(ref, pt0, "num_filter_functions")
    nr, np_ = _both(ref, pt0, "filter_function_name", C.c_char_p)
    ref_names = {nr(i) for i in range(cr())}
    ours = [np_(i) for i in range(cp())]
    assert len(ours) == len(set(ours)) >= 20
    assert set(ours) <= ref_names, set(ours) - ref_names


def test_crop_helpers(ref, pt0):
    """common.h rect helpers (normalize / round / aspect_set / aspect_fit / stretch / offset /
    rotate): bit-identical with the compiled reference on random and degenerate rects."""
    fr, fp_ = _both(ref, pt0, "rect2df_op", C.c_float)
    rng = np.random.default_rng(21)
    rects = [np.array(r, T.f32) for r in ([0, 0, 1920, 1080], [0, 0, 1024, 768], [1920, 1080, 0, 0],
                                         [5, 5, 5, 90], [0, 0, 0, 0], [-3.5, 2.25, 100.75, -40.5])]
    rects += [rng.uniform(-2000, 4000, 4).astype(T.f32) for _ in range(60)]
    for rc in rects:
        for op in range(7):
            for _ in range(6):
                a = rng.uniform(-1.0, 3.0, 5).astype(T.f32)
                if op == 2:
                    a[:4] = rng.uniform(-500, 3000, 4)
                    a[4] = rng.choice([0.0, 0.5, 1.0])
                if op == 1:
                    a[0] = abs(a[0]) * rng.choice([0.0, 1.0, 1.0])
                    a[1] = rng.choice([0.0, 0.3, 1.0])
                if op == 5:
                    a[0] = rng.integers(-9, 10)
                x, y = rc.copy(), rc.copy()
                ra = fr(op, T.fp(x), T.fp(a))
                rb = fp_(op, T.fp(y), T.fp(a))
                assert np.array_equal(x.view(np.uint32), y.view(np.uint32)), (op, rc, a)
                assert np.float32(ra).view(np.uint32) == np.float32(rb).view(np.uint32), (op, rc, a)
    gr, gp = _both(ref, pt0, "rect3_op")
    for kind in range(4):
        for _ in range(40):
            v = (rng.uniform(-100, 100, 6) * (1 if kind in (1, 2) else 1)).astype(T.f32)
            if kind in (0, 3):
                v = np.round(v)
            x, y = v.copy(), v.copy()
            gr(kind, T.fp(x))
            gp(kind, T.fp(y))
            assert np.array_equal(x.view(np.uint32), y.view(np.uint32)), (kind, v)


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


def test_matrix2x2_family(ref, pt0):
    """pl_matrix2x2_* / pl_transform2x2_* / *_apply_rc / pl_transform2x2_bounds and the two
    colour-temperature loci: bit-identical with the compiled reference."""
    fr, fp_ = _both(ref, pt0, "mat2_op")
    rng = np.random.default_rng(31)
    for op in range(14):
        for _ in range(40):
            t = rng.uniform(-3, 3, 6).astype(T.f32)
            u = rng.uniform(-3, 3, 6).astype(T.f32)
            rc = rng.uniform(-500, 2000, 4).astype(T.f32)
            v = rng.uniform(-7, 7, 2).astype(T.f32)
            a = [x.copy() for x in (t, u, rc, v)]
            b = [x.copy() for x in (t, u, rc, v)]
            fr(op, *[T.fp(x) for x in a])
            fp_(op, *[T.fp(x) for x in b])
            for x, y in zip(a, b):
                assert np.array_equal(x.view(np.uint32), y.view(np.uint32)), (op, t, u, rc, v)
    gr, gp = _both(ref, pt0, "mat3_apply_rc")
    for off in (0, 1):
        for _ in range(40):
            m = rng.uniform(-2, 2, 12).astype(T.f32)
            rc = rng.uniform(-1, 2, 6).astype(T.f32)
            x, y = rc.copy(), rc.copy()
            gr(T.fp(m), off, T.fp(x))
            gp(T.fp(m), off, T.fp(y))
            assert np.array_equal(x.view(np.uint32), y.view(np.uint32)), (off, m, rc)
    lr, lp = _both(ref, pt0, "locus_from_temp")
    for kind in (0, 1):
        for temp in (500, 1667, 2222, 2500, 4000, 5000, 6504, 7000, 12000, 25000, 60000):
            x, y = np.zeros(2, T.f32), np.zeros(2, T.f32)
            lr(kind, C.c_float(temp), T.fp(x))
            lp(kind, C.c_float(temp), T.fp(y))
            assert np.array_equal(x.view(np.uint32), y.view(np.uint32)), (kind, temp)
