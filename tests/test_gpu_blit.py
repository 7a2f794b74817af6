"""pl_tex_blit (gpu.h:849) and pl_frame_clear_rgba (renderer.h:675) on the CUDA backend."""
import ctypes as C

import numpy as np
import pytest

import plbtest as T

pytestmark = pytest.mark.gpu


def _blit(gpu, src, dst, src_rc=(0, 0, 0, 0), dst_rc=(0, 0, 0, 0), linear=False):
    s, d = np.array(src_rc, np.int32), np.array(dst_rc, np.int32)
    gpu.lib.shim_tex_blit(C.c_void_p(gpu.ctx), C.c_void_p(src), C.c_void_p(dst), T.ip(s), T.ip(d), int(linear))


@pytest.mark.parametrize("fmt", ["rgba8", "rgba16hf", "r16"])
def test_blit_copies_subrects_and_preserves_the_rest(gpu, fmt):
    dt, comps = {"rgba8": (np.uint8, 4), "rgba16hf": (np.float16, 4), "r16": (np.uint16, 1)}[fmt]
    rng = np.random.default_rng(3)
    w, h = 96, 64
    if dt == np.float16:
        a = rng.random((h, w, comps)).astype(dt)
        b = rng.random((h, w, comps)).astype(dt)
    else:
        a = rng.integers(0, np.iinfo(dt).max, (h, w, comps)).astype(dt)
        b = rng.integers(0, np.iinfo(dt).max, (h, w, comps)).astype(dt)
    src, dst = gpu.tex(fmt, w, h, a), gpu.tex(fmt, w, h, b)
    try:
        _blit(gpu, src, dst, (8, 4, 40, 36), (50, 20, 82, 52))     # 32 x 32 region, same size
        got = gpu.download(dst, fmt, w, h)
        want = b.copy()
        want[20:52, 50:82] = a[4:36, 8:40]
        assert np.array_equal(got.view(np.uint8), want.view(np.uint8))
        _blit(gpu, src, dst)                                        # whole texture
        assert np.array_equal(gpu.download(dst, fmt, w, h).view(np.uint8), a.view(np.uint8))
        _blit(gpu, src, dst, (0, 0, w, h), (w, 0, 0, h))            # horizontal flip
        assert np.array_equal(gpu.download(dst, fmt, w, h).view(np.uint8), a[:, ::-1].copy().view(np.uint8))
    finally:
        gpu.free(src, dst)


def test_blit_scales_with_nearest_sampling(gpu):
    rng = np.random.default_rng(4)
    a = rng.integers(0, 255, (16, 16, 4)).astype(np.uint8)
    src, dst = gpu.tex("rgba8", 16, 16, a), gpu.tex("rgba8", 32, 32)
    try:
        _blit(gpu, src, dst)
        got = gpu.download(dst, "rgba8", 32, 32)
        assert np.array_equal(got, np.repeat(np.repeat(a, 2, 0), 2, 1))
    finally:
        gpu.free(src, dst)


def test_frame_clear_rgba_ycbcr_and_rgb(gpu):
    w, h = 32, 16
    planes = [gpu.tex("r8", w, h), gpu.tex("r8", w // 2, h // 2), gpu.tex("r8", w // 2, h // 2)]
    rgb = gpu.tex("rgba8", w, h)
    csp = T.csp(T.PRIM["BT_709"], T.TRC["BT_1886"])
    try:
        arr = (C.c_void_p * 3)(*planes)
        for rgba, want in (((0, 0, 0, 1), (16, 128, 128)), ((1, 1, 1, 1), (235, 128, 128))):
            col = np.array(rgba, T.f32)
            gpu.lib.shim_frame_clear(C.c_void_p(gpu.ctx), 3, arr, T.SYS["BT_709"], T.LEVELS["LIMITED"], 8,
                                     T.fp(csp), T.fp(col))
            for p, (pw, ph), v in zip(planes, ((w, h), (w // 2, h // 2), (w // 2, h // 2)), want):
                got = gpu.download(p, "r8", pw, ph)
                assert np.all(got == v), (rgba, v, np.unique(got))
        one = (C.c_void_p * 1)(rgb)
        col = np.array((1.0, 0.5, 0.0, 1.0), T.f32)
        gpu.lib.shim_frame_clear(C.c_void_p(gpu.ctx), 1, one, T.SYS["RGB"], T.LEVELS["FULL"], 8, T.fp(csp), T.fp(col))
        got = gpu.download(rgb, "rgba8", w, h)
        # same transfer family (BT.1886 re-used for the sRGB colour) and same primaries: identity
        assert np.all(got[..., 0] == 255) and np.all(np.abs(got[..., 1].astype(int) - 128) <= 1)
        assert np.all(got[..., 2] == 0) and np.all(got[..., 3] == 255)
    finally:
        gpu.free(rgb, *planes)
