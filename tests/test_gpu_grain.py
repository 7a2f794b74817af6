"""pl_shader_film_grain (AV1) through the C ABI against the oracle fed with the LUTs harvested
from the real reference shader (so the host tables AND the pixel arithmetic are covered; the
tables alone are compared in tests/test_grain_tables.py)."""
import ctypes as C
import os

import numpy as np
import pytest

import pipeline_oracle as P
import plbtest as T
from test_grain_tables import HARVEST, harvest, harvest_h274

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not os.path.exists(HARVEST), reason="reference harvester not built")]


def oracle_params(descs, chroma, depth, overlap, keep):
    fx = T.av1_fixture()
    p = T.OrcAv1Grain()
    if chroma:
        uv = np.fromfile(descs[0][3], np.float32).reshape(-1, 2)
        g = [np.ascontiguousarray(uv[:, 0]), np.ascontiguousarray(uv[:, 1])]
        sc = [np.fromfile(descs[2][3], np.float32), np.fromfile(descs[3][3], np.float32)]
        p.grain_w = p.grain_h = 32
        p.sub_x = p.sub_y = 1
    else:
        g = [np.fromfile(descs[0][3], np.float32)]
        sc = [np.fromfile(descs[2][3], np.float32)]
        p.grain_w = p.grain_h = 64
    off = np.fromfile(descs[1][3], np.uint32)
    keep += g + sc + [off]
    for i in range(len(g)):
        p.grain[i] = T.fp(g[i])
        p.scaling[i] = T.fp(sc[i])
        p.is_luma[i] = 0 if chroma else 1
        p.uv_mult[i] = fx.uv_mult[i] / 64.0
        p.uv_mult_luma[i] = fx.uv_mult_luma[i] / 64.0
        # bits = repr.bits.color_depth AFTER pl_color_repr_normalize = the sample depth (16)
        p.uv_offset[i] = float(int(fx.uv_offset[i]) << (16 - 8)) / ((1 << depth) - 1)
    p.offsets = off.ctypes.data_as(C.POINTER(C.c_uint32))
    p.offsets_w, p.offsets_h = descs[1][1], descs[1][2]
    p.num_comps = len(g)
    p.overlap = overlap
    p.chroma_scaling_from_luma = int(fx.chroma_scaling_from_luma)
    p.bit_depth = depth
    p.scale = float(np.float32((1 << 16) / (1 << depth)))  # limited range: 2^16 / 2^depth
    out_scale = np.float32(65536 / 65535.0)  # film_grain_av1.c:866-872: normalised repr, 16 bits
    p.minValue = float(np.float32(16 / 256.0 * float(out_scale)))
    p.maxLuma = float(np.float32(235 / 256.0 * float(out_scale)))
    p.maxChroma = float(np.float32(240 / 256.0 * float(out_scale)))
    p.scaling_shift = fx.scaling_shift
    p.luma_comp = 0
    return p


@pytest.mark.parametrize("overlap", [0, 1])
@pytest.mark.parametrize("chroma", [0, 1])
def test_av1_grain_plane(gpu, oracle, chroma, overlap):
    w, h, seed, depth = 256, 144, 1234, 10
    y, u, v = T.yuv420p10(w, h, seed=9)
    luma_tex = gpu.tex("r16", w, h, y)
    if chroma:
        uv = np.ascontiguousarray(np.stack([u, v], -1))
        tex = gpu.tex("rg16", w // 2, h // 2, uv)
        pw, ph = w // 2, h // 2
    else:
        tex, pw, ph = luma_tex, w, h
    dst = gpu.tex("rgba32f", pw, ph)
    try:
        rc = gpu.lib.shim_film_grain(C.c_void_p(gpu.ctx), C.c_void_p(tex), C.c_void_p(luma_tex),
                                     C.c_void_p(dst), chroma, seed, overlap, depth, 1, 2)
        assert rc == 0
        got = gpu.download(dst, "rgba32f", pw, ph)
    finally:
        gpu.free(dst, luma_tex, *([tex] if chroma else []))
    keep = []
    p = oracle_params(harvest(chroma, w, h, seed, depth), chroma, depth, overlap, keep)
    lf = T.unorm16_to_float(y)[:, :, None]
    plane = T.unorm16_to_float(uv) if chroma else lf
    want = oracle.av1_grain(plane, lf, p)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), np.mean(got != want)
    # the grain must actually have changed the plane
    base = plane[..., 0] * np.float32(64)
    assert np.mean(got[..., 0] != base) > 0.5


@pytest.mark.parametrize("deband", [False, True])
def test_render_with_av1_grain(gpu, deband):
    """Config C4's plane stage at reduced size through pl_render_image: merged chroma planes,
    grain per plane (after debanding when enabled), EWA chroma upscale from the grained FBO."""
    import ctypes as C
    from test_gpu_render import SDR, gpu_render
    gpu.lib.shim_renderer_reset(C.c_void_p(gpu.ctx))
    planes = T.yuv420p10(256, 144, seed=17)
    img = P.Frame(planes, T.SYS["BT_709"], csp=SDR)
    tgt = P.Target(256, 144, "rgba16", sample_depth=16, color_depth=16, csp=SDR)
    got, _ = gpu_render(gpu, img, tgt, upscaler="ewa_lanczos", deband=deband, grain=(1234, 1), frame_index=3)
    want, _ = P.render(img, tgt, upscaler="ewa_lanczos", deband=deband, film_grain=(1234, 1), frame_index=3)
    d = np.abs(got[..., :3].astype(np.int32) - want[..., :3].astype(np.int32))
    assert d.max() == 0, (d.max(), np.mean(d > 0))


# ---- H.274 -----------------------------------------------------------------------------------
H274_FIXTURE = dict(log2_scale=2, present=(1, 0, 0), nvals=(3, 0, 0),
                    intervals={0: [(10, 250, (16, 12, 14))]})
H274_MULTI = dict(log2_scale=3, present=(1, 1, 1), nvals=(3, 2, 1),
                  intervals={0: [(0, 60, (20, 4, 6)), (61, 140, (-14, 9, 3)), (150, 255, (9, 14, 2))],
                             1: [(0, 120, (31, 7, 0)), (121, 255, (12, 11, 0))],
                             2: [(20, 235, (24, 0, 0))]})


def h274_run(gpu, oracle, tex_fmt, texels, comps, cmap, model, seed, depth, sys_, db):
    """Runs pl_shader_film_grain (H.274) on the GPU and the oracle with the same model."""
    h, w = texels.shape[:2]
    lower, upper = np.zeros((3, 8), np.uint8), np.zeros((3, 8), np.uint8)
    values = np.zeros((3, 8, 6), np.int16)
    nint = [0, 0, 0]
    for ch, ivs in model["intervals"].items():
        nint[ch] = len(ivs)
        for i, (lo, hi, vals) in enumerate(ivs):
            lower[ch, i], upper[ch, i] = lo, hi
            values[ch, i, :len(vals)] = vals
    ia = lambda v: (C.c_int * len(v))(*v)  # noqa: E731
    tex = gpu.tex(tex_fmt, w, h, texels)
    dst = gpu.tex("rgba32f", w, h)
    try:
        rc = gpu.lib.shim_film_grain_h274(
            C.c_void_p(gpu.ctx), C.c_void_p(tex), C.c_void_p(dst), comps, ia(list(cmap) + [0] * (4 - len(cmap))),
            seed, depth, 1, sys_, model["log2_scale"], ia(model["present"]), ia(nint), ia(model["nvals"]),
            lower.ctypes.data_as(C.c_void_p), upper.ctypes.data_as(C.c_void_p),
            values.ctypes.data_as(C.c_void_p), 0)
        assert rc == 0, rc
        got = gpu.download(dst, "rgba32f", w, h)
    finally:
        gpu.free(dst, tex)

    # the same uniforms as film_grain_h274.c:190-282 computes them
    p = T.OrcH274Grain()
    keep = [np.ascontiguousarray(db, np.float32)]
    p.db, p.db_w, p.seed = T.fp(keep[0]), 832, seed
    p.scale = float(np.float32((1 << 16) / (1 << depth)))  # limited range, 16-bit samples
    p.scale_factor = float(np.float32(1.0 / (1 << (model["log2_scale"] + 6))))
    ycc = sys_ not in (T.SYS["RGB"], T.SYS["XYZ"], T.SYS["UNKNOWN"])
    n = 0
    for idx in range(comps):
        ch = cmap[idx] if ycc else {0: 2, 1: 0, 2: 1}[cmap[idx]]  # film_grain.h:31-70
        if not model["present"][ch]:
            continue
        lo = np.array([iv[0] / 255.0 for iv in model["intervals"][ch]], np.float32)
        hi = np.array([iv[1] / 255.0 for iv in model["intervals"][ch]], np.float32)
        val = []
        for _, _, vals in model["intervals"][ch]:
            nv = model["nvals"][ch]
            hh = vals[1] if nv > 1 else 8
            vv = vals[2] if nv > 2 else hh
            hh, vv = min(max(hh, 2), 14) - 2, min(max(vv, 2), 14) - 2
            sc = vals[0] >> 1 if (ch > 0 and ycc) else vals[0]
            val.append(((sc & 0xFFFF) << 16) | hh << 8 | vv)
        val = np.array(val, np.uint32)
        keep += [lo, hi, val]
        p.channel[n], p.num_intervals[n] = ch, len(val)
        p.lower[n], p.upper[n] = T.fp(lo), T.fp(hi)
        p.value[n] = val.ctypes.data_as(C.POINTER(C.c_uint32))
        n += 1
    p.num_entries = n
    plane = T.unorm16_to_float(texels if texels.ndim == 3 else texels[:, :, None])
    want = oracle.h274_grain(plane, p)
    return got, want, plane, float(p.scale)


@pytest.fixture(scope="module")
def h274_db():
    db, _ = harvest_h274()
    return db


@pytest.mark.parametrize("w,h", [(256, 144), (250, 131)])
def test_h274_grain_luma_fixture(gpu, oracle, h274_db, w, h):
    """The reference's own fixture on a luma plane (gpu_tests.c:787-835 runs the same call);
    the second size leaves partial 8x8 groups at the right / bottom edge."""
    y = T.yuv420p10(256, 144, seed=11)[0][:h, :w]
    got, want, plane, scale = h274_run(gpu, oracle, "r16", np.ascontiguousarray(y), 1, (0,), H274_FIXTURE,
                                       1234, 10, T.SYS["BT_709"], h274_db)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), np.mean(got != want)
    assert np.mean(got[..., 0] != plane[..., 0] * np.float32(scale)) > 0.5


@pytest.mark.parametrize("sysn", ["BT_2020_NC", "RGB"])
def test_h274_grain_three_components_intervals(gpu, oracle, h274_db, sysn):
    """All three components, several intensity intervals (gaps select no grain), negative
    scale, chroma halving for YCbCr-like systems, the RGB -> (Cr, Y, Cb) channel map."""
    rng = np.random.default_rng(5)
    yy, xx = np.mgrid[0:96, 0:160]
    base = ((xx * 400 + yy * 150) % 65536).astype(np.float64)
    tex = np.clip(base[..., None] + rng.integers(-900, 900, (96, 160, 4)), 0, 65535).astype(np.uint16)
    got, want, _, _ = h274_run(gpu, oracle, "rgba16", tex, 3, (0, 1, 2), H274_MULTI, 77, 10,
                               T.SYS[sysn], h274_db)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), np.mean(got != want)


def test_h274_not_needed(gpu):
    """pl_needs_film_grain: model_id != 0 or no present component -> nothing to do
    (film_grain_h274.c:127-143); a chroma-only plane with the luma-only fixture."""
    ia = lambda v: (C.c_int * len(v))(*v)  # noqa: E731
    z8, z16 = np.zeros((3, 8), np.uint8), np.zeros((3, 8, 6), np.int16)
    tex = gpu.tex("rg16", 64, 64, np.zeros((64, 64, 2), np.uint16))
    dst = gpu.tex("rgba32f", 64, 64)
    try:
        rc = gpu.lib.shim_film_grain_h274(
            C.c_void_p(gpu.ctx), C.c_void_p(tex), C.c_void_p(dst), 2, ia([1, 2, 0, 0]), 1, 10, 1, 2, 2,
            ia([1, 0, 0]), ia([1, 0, 0]), ia([3, 0, 0]), z8.ctypes.data_as(C.c_void_p),
            z8.ctypes.data_as(C.c_void_p), z16.ctypes.data_as(C.c_void_p), 0)
        assert rc == 1
    finally:
        gpu.free(dst, tex)


@pytest.mark.parametrize("deband", [False, True])
def test_render_with_h274_grain(gpu, deband):
    """pl_render_image with the reference's H.274 fixture in pl_frame.film_grain: a grain pass on
    the luma plane only (after debanding when enabled), none on chroma."""
    from test_gpu_render import SDR, gpu_render
    gpu.lib.shim_renderer_reset(C.c_void_p(gpu.ctx))
    planes = T.yuv420p10(256, 144, seed=19)
    img = P.Frame(planes, T.SYS["BT_709"], csp=SDR)
    tgt = P.Target(256, 144, "rgba16", sample_depth=16, color_depth=16, csp=SDR)
    got, _ = gpu_render(gpu, img, tgt, upscaler="ewa_lanczos", deband=deband, grain=(4321, "h274"), frame_index=2)
    want, _ = P.render(img, tgt, upscaler="ewa_lanczos", deband=deband, film_grain=(4321, "h274"), frame_index=2)
    d = np.abs(got[..., :3].astype(np.int32) - want[..., :3].astype(np.int32))
    assert d.max() == 0, (d.max(), np.mean(d > 0))
    plain, _ = P.render(img, tgt, upscaler="ewa_lanczos", deband=deband, frame_index=2)
    assert np.mean(plain[..., :3] != want[..., :3]) > 0.3
