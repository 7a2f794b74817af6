"""Frame ingest through utils/upload.h on the GPU: pl_upload_plane (host pointer, pl_buf, row
padding, foreign-endian samples via the byte-swap kernel), pl_recreate_plane, and a semi-planar
P010 frame rendered by pl_render_image against the CPU oracle."""
import ctypes as C

import numpy as np
import pytest

import pipeline_oracle as P
import plbtest as T
from test_upload import MASKS, UNORM, FLOAT, Upload

pytestmark = pytest.mark.gpu


def upload(gpu, typ, layout, pixel_stride, rows, *, swapped=0, via_buf=0, recreate=False, w=None):
    """rows: (h, row_stride) uint8 array of host bytes. Returns (tex, info, raw texel bytes)."""
    h, row_stride = rows.shape
    w = w or row_stride // pixel_stride
    info = np.zeros(8, np.int32)
    f = gpu.lib.shim_upload_plane
    f.restype = C.c_void_p
    lay = np.ascontiguousarray(layout, np.int32)
    tex = f(C.c_void_p(gpu.ctx), typ, lay.ctypes.data_as(C.c_void_p), pixel_stride, row_stride, swapped,
            w, h, None if recreate else rows.ctypes.data_as(C.c_void_p), C.c_size_t(rows.size), via_buf,
            info.ctypes.data_as(C.c_void_p))
    if not tex:
        return None, info, None
    raw = np.zeros((h, w * pixel_stride), np.uint8)
    assert gpu.lib.shim_tex_download(C.c_void_p(gpu.ctx), C.c_void_p(tex), raw.ctypes.data_as(C.c_void_p)) == 0
    return tex, info, raw


@pytest.mark.parametrize("via_buf", [0, 1])
@pytest.mark.parametrize("name,stride,cmap", [("y8", 1, None), ("nv12_uv", 2, (1, 2)), ("rgba8", 4, None),
                                              ("bgra8", 4, None), ("rgba16", 8, None)])
def test_upload_plane_copies_rows_with_padding(gpu, name, stride, cmap, via_buf):
    up = Upload(T.shim(), "plt_")
    lay = up.from_mask(MASKS[name])
    if cmap:
        lay[8:8 + len(cmap)] = cmap
    w, h, pad = 50, 9, 6 * stride
    rng = np.random.default_rng(1)
    rows = rng.integers(0, 256, (h, w * stride + pad), dtype=np.uint8)
    tex, info, raw = upload(gpu, UNORM, lay, stride, rows, via_buf=via_buf, w=w)
    assert tex
    try:
        assert np.array_equal(raw, rows[:, :w * stride])
        ncomp = sum(1 for m in MASKS[name] if m)
        assert info[0] == ncomp and info[5] >= ncomp and info[6] == stride
        want_map = {"bgra8": [2, 1, 0, 3], "nv12_uv": [1, 2, -1, -1]}.get(name, list(range(ncomp)) + [-1] * (4 - ncomp))
        assert list(info[1:5]) == want_map
    finally:
        gpu.free(tex)


@pytest.mark.parametrize("via_buf", [0, 1, 2])
@pytest.mark.parametrize("name,stride", [("p010_y", 2), ("p010_uv", 4), ("rgba16", 8)])
def test_upload_plane_foreign_endian(gpu, name, stride, via_buf):
    """Big-endian 16-bit samples: swapped on the device (pl_buf_copy_swap semantics), from host
    memory, from a plain buffer (copied first) and from a storable buffer (swapped out of it)."""
    up = Upload(T.shim(), "plt_")
    ok, lay, bits = up.align(up.from_mask(MASKS[name]), stride)
    assert ok
    w, h = 37, 7
    rng = np.random.default_rng(2)
    native = rng.integers(0, 65536, (h, w * stride // 2), dtype=np.uint16)
    rows = native.byteswap().view(np.uint8).reshape(h, -1)
    tex, info, raw = upload(gpu, UNORM, lay, stride, np.ascontiguousarray(rows), swapped=1, via_buf=via_buf, w=w)
    assert tex
    try:
        assert np.array_equal(raw.view(np.uint16), native)
    finally:
        gpu.free(tex)


def test_recreate_plane_makes_a_render_target(gpu):
    up = Upload(T.shim(), "plt_")
    lay = up.from_mask(MASKS["rgba16"])
    tex, info, raw = upload(gpu, FLOAT, lay, 8, np.zeros((8, 8 * 16), np.uint8), recreate=True, w=16)
    assert tex
    try:
        assert info[0] == 4 and list(info[1:5]) == [0, 1, 2, 3] and info[7] == 1
    finally:
        gpu.free(tex)
    # a layout no format provides
    tex, _, _ = upload(gpu, UNORM, up.from_mask(MASKS["rgb565"]), 2, np.zeros((4, 32), np.uint8))
    assert tex is None


def test_render_p010_semi_planar_frame(gpu):
    """A P010 frame (Y: r16, CbCr: rg16, 10 significant bits in the top of each 16-bit sample,
    bit_shift 6 from pl_plane_data_align) ingested with pl_upload_plane and rendered with
    pl_render_image equals the oracle's render of the same code values."""
    from test_gpu_render import SDR
    up = Upload(T.shim(), "plt_")
    gpu.lib.shim_renderer_reset(C.c_void_p(gpu.ctx))
    Y, U, V = T.yuv420p10(256, 144, seed=23)
    Y6, U6, V6 = (Y << 6).astype(np.uint16), (U << 6).astype(np.uint16), (V << 6).astype(np.uint16)
    ok, lay_y, bits = up.align(up.from_mask(MASKS["p010_y"]), 2)
    ok2, lay_uv, bits2 = up.align(up.from_mask(MASKS["p010_uv"]), 4)
    assert ok and ok2 and list(bits) == [16, 10, 6] and list(bits2) == [16, 10, 6]
    lay_uv[8:10] = (1, 2)
    uv = np.ascontiguousarray(np.stack([U6, V6], -1))
    ty, _, _ = upload(gpu, UNORM, lay_y, 2, np.ascontiguousarray(Y6).view(np.uint8).reshape(144, -1), w=256)
    tuv, info, _ = upload(gpu, UNORM, lay_uv, 4, uv.view(np.uint8).reshape(72, -1), w=128)
    assert ty and tuv and list(info[:3]) == [2, 1, 2]
    tgt = P.Target(256, 144, "rgba16", sample_depth=16, color_depth=16, csp=SDR)
    dst = gpu.tex(tgt.fmt, tgt.w, tgt.h)
    try:
        img = P.Frame((Y6, U6, V6), T.SYS["BT_709"], csp=SDR, sample_depth=16, color_depth=10, bit_shift=6)
        planes = (C.c_void_p * 3)(ty, tuv, None)
        iarr = np.concatenate([[img.sys, img.levels, *img.bits], img.csp]).astype(T.f32)
        tarr = np.concatenate([[tgt.sys, tgt.levels, *tgt.bits], tgt.csp]).astype(T.f32)
        opts = np.zeros(16, np.int32)
        opts[5] = P.CHROMA_ENUM[img.chroma_loc]
        peak_out = np.zeros(2, T.f32)
        r = gpu.lib.shim_render(C.c_void_p(gpu.ctx), 2, planes, C.c_void_p(dst), T.fp(iarr), T.fp(tarr),
                                T.ip(opts), b"ewa_lanczos", b"", b"", T.fp(peak_out))
        assert r == 0
        got = gpu.download(dst, tgt.fmt, tgt.w, tgt.h)
    finally:
        gpu.free(dst, ty, tuv)
    want, _ = P.render(img, tgt, upscaler="ewa_lanczos")
    d = np.abs(got[..., :3].astype(np.int32) - want[..., :3].astype(np.int32))
    assert d.max() == 0, (d.max(), np.mean(d > 0))
