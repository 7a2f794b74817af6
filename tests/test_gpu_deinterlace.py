"""pl_shader_deinterlace (shaders/deinterlacing.c) through the C ABI against the oracle: weave,
bob, yadif (with and without the spatial check) and bwdif, both fields, both field orders, with
and without neighbouring frames, one- and multi-component textures; and inside pl_render_image
(interlaced HDR10 4:2:0 frame, temporal references, renderer.c:1591-1612). Bit-exact."""
import ctypes as C

import numpy as np
import pytest

import pipeline_oracle as P
import plbtest as T
from test_gpu_render import HDR10, SDR, gpu_render

pytestmark = pytest.mark.gpu
ALGO = {"weave": 0, "bob": 1, "yadif": 2, "bwdif": 3}


def run(gpu, fmt, cur, prev, nxt, field, first, algo, skip=False, mask=0):
    h, w = cur.shape[:2]
    texs = [gpu.tex(fmt, w, h, a) if a is not None else None for a in (cur, prev, nxt)]
    dst = gpu.tex("rgba32f", w, h)
    try:
        r = gpu.lib.shim_deinterlace(C.c_void_p(gpu.ctx), C.c_void_p(texs[0]),
                                     C.c_void_p(texs[1]) if texs[1] else None,
                                     C.c_void_p(texs[2]) if texs[2] else None, C.c_void_p(dst),
                                     field, first, algo, int(skip), mask)
        assert r == 0
        return gpu.download(dst, "rgba32f", w, h)
    finally:
        gpu.free(dst, *[t for t in texs if t])


@pytest.mark.parametrize("algo", ["weave", "bob", "yadif", "bwdif"])
@pytest.mark.parametrize("field,first", [(1, 0), (2, 0), (1, 2), (2, 2)])
@pytest.mark.parametrize("refs", ["both", "none", "prev"])
def test_deinterlace_shader(gpu, oracle, algo, field, first, refs):
    rng = np.random.default_rng(ALGO[algo] * 10 + field)
    shape = (37, 53, 1)
    cur, prev, nxt = (rng.random(shape, dtype=np.float32) for _ in range(3))
    if refs == "none":
        prev = nxt = None
    elif refs == "prev":
        nxt = None
    got = run(gpu, "r32f", cur, prev, nxt, field, first, ALGO[algo])
    want = oracle.deinterlace(cur, prev, nxt, field=field, first_field=first, algo=ALGO[algo])
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), float(np.abs(got - want).max())


def test_deinterlace_multi_component_and_options(gpu, oracle):
    rng = np.random.default_rng(3)
    cur, prev, nxt = (rng.random((40, 64, 4), dtype=np.float32) for _ in range(3))
    for algo, skip, mask in (("yadif", True, 0), ("yadif", False, 0b0101), ("bwdif", False, 0b0111)):
        got = run(gpu, "rgba32f", cur, prev, nxt, 1, 1, ALGO[algo], skip, mask)
        want = oracle.deinterlace(cur, prev, nxt, field=1, first_field=1, algo=ALGO[algo],
                                  skip_spatial_check=skip, comp_mask=mask)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # field NONE: passthrough
    got = run(gpu, "rgba32f", cur, None, None, 0, 0, ALGO["yadif"])
    assert np.array_equal(got, cur)


@pytest.mark.parametrize("algo,field", [("yadif", 1), ("bwdif", 2), ("bob", 1)])
def test_render_interlaced_frame(gpu, algo, field):
    """C3-style render of an interlaced frame with temporal neighbours."""
    gpu.lib.shim_renderer_reset(C.c_void_p(gpu.ctx))
    imgs = [P.Frame(T.yuv420p10(160, 92, seed=70 + i, kind="noise" if i == 1 else "rings"),
                    T.SYS["BT_2020_NC"], csp=HDR10) for i in range(3)]
    tgt = P.Target(160, 92, "rgba16", sample_depth=16, color_depth=16, csp=SDR)
    kw = dict(plane_upscaler="ewa_lanczos", deband=True, peak=2, dither=0, cmap_hq=True, frame_index=3)
    ptex = [gpu.tex("r16", p.shape[1], p.shape[0], p) for p in imgs[0].planes]
    ntex = [gpu.tex("r16", p.shape[1], p.shape[0], p) for p in imgs[2].planes]
    try:
        gpu.lib.shim_set_deinterlace(C.c_void_p(gpu.ctx), field, 1, ALGO[algo], 0,
                                     (C.c_void_p * 3)(*ptex), (C.c_void_p * 3)(*ntex))
        got, _ = gpu_render(gpu, imgs[1], tgt, **kw)
    finally:
        gpu.free(*ptex, *ntex)
    want, _ = P.render(imgs[1], tgt, deinterlace=(field, 1, ALGO[algo], imgs[0], imgs[2]), **kw)
    d = np.abs(got[..., :3].astype(np.int64) - want[..., :3].astype(np.int64))
    assert d.max() == 0, (int(d.max()), float(np.mean(d > 0)))
    plain, _ = P.render(imgs[1], tgt, **kw)
    assert np.any(plain != want)
