"""pl_shader_film_grain (AV1) through the C ABI against the oracle fed with the LUTs harvested
from the real reference shader (so the host tables AND the pixel arithmetic are covered; the
tables alone are compared in tests/test_grain_tables.py)."""
import ctypes as C
import os

import numpy as np
import pytest

import pipeline_oracle as P
import plbtest as T
from test_grain_tables import HARVEST, harvest

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
