"""Custom LUTs on the GPU (pl_shader_custom_lut, shaders/lut.c:212-280; renderer hooks
renderer.c:1919-1946, :2198-2247, :2274-2276, :2776-2785) against the CPU oracle: the shader on
its own (1-D, cubic and non-cubic 3-D tables, shaper matrices) and inside pl_render_image
(main LUT native / conversion, image LUT normalized, target LUT native). Bit-exact."""
import ctypes as C

import numpy as np
import pytest

import pipeline_oracle as P
import plbtest as T
from test_gpu_render import HDR10, SDR, gpu_render
from test_lut import identity_lut

pytestmark = pytest.mark.gpu


def run_lut(gpu, px, data, size, shaper_in=None, shaper_out=None, sig=1):
    h, w = px.shape[:2]
    src = gpu.tex("rgba32f", w, h, px)
    dst = gpu.tex("rgba32f", w, h)
    try:
        d = np.ascontiguousarray(data, np.float32)
        si = None if shaper_in is None else T.fp(np.ascontiguousarray(shaper_in, np.float32))
        so = None if shaper_out is None else T.fp(np.ascontiguousarray(shaper_out, np.float32))
        r = gpu.lib.shim_custom_lut(C.c_void_p(gpu.ctx), C.c_void_p(src), C.c_void_p(dst), T.fp(d),
                                    (C.c_int * 3)(*size), si, so, C.c_ulonglong(sig))
        assert r == 0
        return gpu.download(dst, "rgba32f", w, h)
    finally:
        gpu.free(src, dst)


@pytest.mark.parametrize("size,shapers", [((17, 17, 17), False), ((33, 33, 33), True), ((9, 5, 7), False),
                                          ((2, 2, 2), False), ((64, 0, 0), False), ((1024, 0, 0), True)])
def test_custom_lut_shader(gpu, size, shapers):
    rng = np.random.default_rng(sum(size))
    n = size[0] * (size[1] or 1) * (size[2] or 1)
    tab = rng.uniform(-0.1, 1.1, (n, 3)).astype(np.float32)
    px = rng.uniform(-0.2, 1.2, (97, 131, 4)).astype(np.float32)
    # exact lattice points and ties between the fractional parts (all six tetrahedron tests)
    px[0, :, :3] = (rng.integers(0, 4, (131, 3)) / np.float32(4)).astype(np.float32)
    px[1, :, 0] = px[1, :, 1]
    px[2, :, 1] = px[2, :, 2]
    px[3, :, :3] = px[3, :, :1]
    si = so = None
    if shapers:
        si = (np.eye(3) + rng.uniform(-0.1, 0.1, (3, 3))).astype(np.float32).ravel()
        so = (np.eye(3) + rng.uniform(-0.1, 0.1, (3, 3))).astype(np.float32).ravel()
    got = run_lut(gpu, px, tab, size, si, so, sig=1000 + sum(size))
    want = T.oracle().custom_lut(px, tab, size, si, so)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), \
        float(np.abs(got - want).max())


def perturbed_identity(n, seed, amp=0.02):
    rng = np.random.default_rng(seed)
    return np.clip(identity_lut(n) + rng.uniform(-amp, amp, (n ** 3, 3)), 0, 1).astype(np.float32)


@pytest.mark.parametrize("which,kind", [("params", "native"), ("params", "conversion"),
                                        ("image", "normalized"), ("target", "native")])
def test_render_with_custom_lut(gpu, which, kind):
    """C3-style frame (HDR10 4:2:0 -> BT.709 rgba16) with one custom 3-D LUT in the pipeline."""
    gpu.lib.shim_renderer_reset(C.c_void_p(gpu.ctx))
    planes = T.yuv420p10(160, 90, seed=31, kind="noise")
    img = P.Frame(planes, T.SYS["BT_2020_NC"], csp=HDR10)
    tgt = P.Target(160, 90, "rgba16", sample_depth=16, color_depth=16, csp=SDR)
    luts = {which: (kind, perturbed_identity(17, 7), (17, 17, 17))}
    kw = dict(plane_upscaler="ewa_lanczos", deband=True, peak=2, dither=0, cmap_hq=True, frame_index=1,
              luts=luts)
    got, _ = gpu_render(gpu, img, tgt, **kw)
    want, _ = P.render(img, tgt, **kw)
    d = np.abs(got[..., :3].astype(np.int64) - want[..., :3].astype(np.int64))
    # stated tolerance +-1 LSB16; achieved: bit-exact
    assert d.max() == 0, (int(d.max()), float(np.mean(d > 0)))
    # and the LUT really took part
    base, _ = P.render(img, tgt, **dict(kw, luts=None))
    assert np.any(base != want)


def test_render_with_normalized_main_lut_identity(gpu):
    """PL_LUT_NORMALIZED main LUT (renderer.c:2213-2239): the image is linearised, mapped to the
    LUT's input space, scaled to the nominal peak, sent through the LUT, and mapped back before
    the main colour conversion runs on the linear image. With an identity table on an SDR frame
    the whole detour must reproduce the plain render up to the rounding of the extra
    linearise / delinearise round trip (<= 2 LSB16; no oracle restatement of this flow)."""
    gpu.lib.shim_renderer_reset(C.c_void_p(gpu.ctx))
    planes = T.yuv420p10(160, 90, seed=41, kind="rings")
    img = P.Frame(planes, T.SYS["BT_709"], csp=SDR)
    tgt = P.Target(160, 90, "rgba16", sample_depth=16, color_depth=16, csp=SDR)
    kw = dict(plane_upscaler="ewa_lanczos", dither=0, frame_index=1)
    base, _ = gpu_render(gpu, img, tgt, **kw)
    gpu.lib.shim_renderer_reset(C.c_void_p(gpu.ctx))
    got, _ = gpu_render(gpu, img, tgt, luts={"params": ("normalized", identity_lut(33), (33, 33, 33))}, **kw)
    d = np.abs(got[..., :3].astype(np.int64) - base[..., :3].astype(np.int64))
    assert d.max() <= 2, (int(d.max()), float(np.mean(d > 0)))
