"""CUDA samplers (through the public C ABI) against the CPU oracle on identical inputs.
Bars: fp32 targets must match the oracle to float rounding (the kernels follow the oracle's
operation order, so we demand <= 1 fp32 ULP and report how many texels are bit-identical);
fp16/unorm16 targets must be within 1 ULP of the 16-bit output (north_star tolerance)."""
import numpy as np
import pytest

import plbtest as T

pytestmark = pytest.mark.gpu


def rand_tex(h, w, c, seed=0, dtype=np.float16):
    rng = np.random.default_rng(seed)
    return rng.random((h, w, c), dtype=np.float32).astype(dtype)


def max_ulp32(a, b):
    ia = a.view(np.int32).astype(np.int64)
    ib = b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia)
    ib = np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    return np.abs(ia - ib).max()


@pytest.mark.parametrize("filt", ["ewa_lanczos", "ewa_lanczossharp", "ewa_hann", "ewa_robidoux"])
@pytest.mark.parametrize("comps", [1, 2, 3, 4])
def test_polar_2x_fast_path_is_bit_exact_with_exact_coordinate_oracle(gpu, oracle, filt, comps):
    fmt = {1: "r16hf", 2: "rg16hf", 3: "rgba16hf", 4: "rgba16hf"}[comps]
    tc = 4 if comps >= 3 else comps
    src = rand_tex(70, 123, tc, seed=comps)
    f = T.t0().filter_generate(filt)
    got = gpu.polar(src, fmt, 246, 140, filt, comps=comps, out_fmt="rgba32f")
    want = oracle.sample_polar(src.astype(np.float32), 246, 140, f, comp_mask=(1 << comps) - 1,
                               coord_mode=T.COORD_EXACT)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_polar_2x_fp16_output_and_literal_oracle_deviation(gpu, oracle):
    """Config C2 main pass at small size: rgba16hf in/out. Against the exact-coordinate
    oracle the result is bit-exact; against the literal fp32 restatement of dispatch.c's
    attribute emulation the deviation is reported (sub-texel drift of the reference's own
    coordinate rounding)."""
    src = rand_tex(135, 240, 4, seed=3)
    f = T.t0().filter_generate("ewa_lanczos")
    got = gpu.polar(src, "rgba16hf", 480, 270, "ewa_lanczos", comps=3, out_fmt="rgba16hf")
    exact = oracle.store_half(oracle.sample_polar(src.astype(np.float32), 480, 270, f, comp_mask=7,
                                                  coord_mode=T.COORD_EXACT))
    assert T.ulp_diff_half(got, exact).max() == 0
    lit = oracle.sample_polar(src.astype(np.float32), 480, 270, f, comp_mask=7,
                              coord_mode=T.COORD_LITERAL)
    dev = np.abs(got.astype(np.float32) - lit).max()
    assert dev < 2e-3  # sanity: same picture


@pytest.mark.parametrize("case", [
    dict(filt="ewa_lanczos", sw=97, sh=61, dw=233, dh=130, comps=3),          # odd ratio
    dict(filt="ewa_lanczos", sw=128, sh=72, dw=90, dh=50, comps=4),           # downscale (widening)
    dict(filt="ewa_lanczos4sharpest", sw=80, sh=60, dw=160, dh=120, comps=3), # anti-ringing
    dict(filt="ewa_lanczossharp", sw=64, sh=48, dw=64, dh=48, comps=2),       # 1:1
    dict(filt="ewa_lanczos", sw=64, sh=48, dw=100, dh=75, comps=3, rect=(3.25, 2.5, 60.0, 44.75)),
])
def test_polar_generic_kernel_matches_literal_oracle(gpu_literal, oracle, case):
    sw, sh, dw, dh, comps = (case[k] for k in ("sw", "sh", "dw", "dh", "comps"))
    tc = 4 if comps >= 3 else comps
    fmt = {1: "r32f", 2: "rg32f", 4: "rgba32f"}[tc]
    src = rand_tex(sh, sw, tc, seed=7, dtype=np.float32)
    rect = case.get("rect", (0, 0, 0, 0))
    orect = rect if any(rect) else None
    rx = dw / ((rect[2] - rect[0]) if any(rect) else sw)
    ry = dh / ((rect[3] - rect[1]) if any(rect) else sh)
    blur = max(1.0 / min(rx, ry), 1.0)
    base = T.t0().filter_generate(case["filt"])
    f = T.t0().filter_generate(case["filt"], blur=(base["blur"] or 1.0) * blur)
    got = gpu_literal.polar(src, fmt, dw, dh, case["filt"], comps=comps, out_fmt="rgba32f", rect=rect)
    want = oracle.sample_polar(src, dw, dh, f, comp_mask=(1 << comps) - 1, rect=orect,
                               coord_mode=T.COORD_LITERAL)
    assert np.isfinite(got).all()
    assert max_ulp32(got, want) <= 1
    ident = np.mean(got.view(np.uint32) == want.view(np.uint32))
    assert ident > 0.999, ident


@pytest.mark.parametrize("dw,dh,comps,rect", [(171, 96, 3, (0, 0, 0, 0)), (90, 50, 4, (0, 0, 0, 0)),
                                              (150, 113, 3, (2.5, 1.25, 120.0, 70.75))])
def test_polar_generic_rgba16hf_source(gpu, gpu_literal, oracle, dw, dh, comps, rect):
    """The image-FBO case of the generic kernel (rgba16hf source, clamp addressing, no
    anti-ringing: format-specialised path) at non-2x ratios, both coordinate modes."""
    sw, sh = 128, 72
    src = rand_tex(sh, sw, 4, seed=9)   # float16 texels
    orect = rect if any(rect) else None
    rx = dw / ((rect[2] - rect[0]) if any(rect) else sw)
    ry = dh / ((rect[3] - rect[1]) if any(rect) else sh)
    blur = max(1.0 / min(rx, ry), 1.0)
    base = T.t0().filter_generate("ewa_lanczos")
    f = T.t0().filter_generate("ewa_lanczos", blur=(base["blur"] or 1.0) * blur)
    for g, mode in ((gpu_literal, T.COORD_LITERAL), (gpu, T.COORD_EXACT)):
        got = g.polar(src, "rgba16hf", dw, dh, "ewa_lanczos", comps=comps, out_fmt="rgba32f", rect=rect)
        want = oracle.sample_polar(src.astype(np.float32), dw, dh, f, comp_mask=(1 << comps) - 1,
                                   rect=orect, coord_mode=mode)
        assert np.isfinite(got).all()
        assert max_ulp32(got, want) <= 1
        assert np.mean(got.view(np.uint32) == want.view(np.uint32)) > 0.999


def test_polar_unorm16_source_and_target(gpu_literal, oracle):
    rng = np.random.default_rng(11)
    src = rng.integers(0, 65536, (40, 50, 2), dtype=np.uint16)
    f = T.t0().filter_generate("ewa_lanczos")
    got = gpu_literal.polar(src, "rg16", 100, 80, "ewa_lanczos", comps=2, out_fmt="rgba16")
    want = oracle.store_unorm(oracle.sample_polar(T.unorm16_to_float(src), 100, 80, f, comp_mask=3,
                                                  coord_mode=T.COORD_LITERAL), 16)
    assert np.abs(got.astype(np.int32) - want.astype(np.int32)).max() <= 1


@pytest.mark.parametrize("kind,linear", [("nearest", False), ("bilinear", True)])
def test_direct_sampling(gpu_literal, oracle, kind, linear):
    src = rand_tex(33, 47, 4, seed=5, dtype=np.float32)
    for (dw, dh, rect) in [(47, 33, (0, 0, 0, 0)), (94, 50, (0, 0, 0, 0)), (30, 20, (2.5, 1.0, 40.0, 30.5))]:
        got = gpu_literal.simple(kind, src, "rgba32f", dw, dh, rect=rect)
        want = oracle.sample_direct(src, dw, dh, linear=linear, rect=rect if any(rect) else None,
                                    coord_mode=T.COORD_LITERAL)
        assert max_ulp32(got, want) <= 1


@pytest.mark.parametrize("comps,fmt", [(1, "r16"), (2, "rg16"), (3, "rgba16")])
def test_deband(gpu, oracle, comps, fmt):
    rng = np.random.default_rng(comps)
    tc = 4 if comps == 3 else comps
    # smooth gradient + small noise in 10-bit code values: both branches of the threshold occur
    base = np.linspace(200, 260, 96)[None, :, None] + np.zeros((64, 96, tc))
    src = np.clip(base + rng.integers(-3, 4, base.shape), 0, 1023).astype(np.uint16)
    scale = np.float32(65535.0 / 1023.0)  # what pl_color_repr_normalize gives for 10-in-16 full range
    for index in (0, 7):
        gpu.set_frame_index(index)
        got = gpu.simple("deband", src, fmt, 96, 64, comps=comps, scale=float(scale),
                         sargs_extra=[1, 3.0, 16.0, 4.0, 0.1, 0.5, 0.5])
        want = oracle.deband(T.unorm16_to_float(src), 1, 3.0, 16.0, 4.0, (0.1, 0.5, 0.5), scale,
                             comp_mask=(1 << comps) - 1, frame_index=index)
        diff = np.abs(got - want)[..., :comps]
        # sin/cos of the random angle come from different libms: a tap can land on the other
        # side of a texel boundary for a handful of pixels
        frac_bad = np.mean(diff.max(-1) > 1e-6)
        assert frac_bad < 2e-3, frac_bad
        assert np.median(diff) == 0
    gpu.set_frame_index(0)


def test_deband_zero_iterations_is_identity(gpu, oracle):
    """src/tests/gpu_tests.c:738-747: deband with 0 iterations and no grain returns the input"""
    src = rand_tex(32, 48, 4, seed=2, dtype=np.float32)
    got = gpu.simple("deband", src, "rgba32f", 48, 32, comps=3, sargs_extra=[0, 0.0, 0.0, 0.0, 0, 0, 0])
    assert np.abs(got[..., :3] - src[..., :3]).max() <= 1e-6


@pytest.mark.parametrize("ox,oy", [(0.25, 0.0), (0.0, 0.25), (0.25, 0.25)])
@pytest.mark.parametrize("comps", [1, 2])
def test_polar_2x_sited_chroma_phases(gpu, oracle, ox, oy, comps):
    """4:2:0 chroma with left / top-left siting: source rect offset by a quarter texel, the
    phase set becomes {0, 1/2}; the tabulated-phase kernel must still equal the oracle."""
    src = rand_tex(54, 96, comps, seed=21)
    fmt = {1: "r16hf", 2: "rg16hf"}[comps]
    f = T.t0().filter_generate("ewa_lanczos")
    rect = (ox, oy, 96 + ox, 54 + oy)
    got = gpu.polar(src, fmt, 192, 108, "ewa_lanczos", comps=comps, out_fmt="rgba32f", rect=rect)
    want = oracle.sample_polar(src.astype(np.float32), 192, 108, f, comp_mask=(1 << comps) - 1,
                               rect=rect, coord_mode=T.COORD_EXACT)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


# ---- separable (ortho) scaling: pl_shader_sample_ortho2, one 1-D pass per call -----------------
ORTHO_FILTERS = ["lanczos", "gaussian", "spline36", "mitchell", "bilinear", "hermite", "ginseng"]


def _ortho_case(gpu, oracle, name, src, dw, dh, comps, antiring=-1.0, fmt=None):
    h, w = src.shape[:2]
    pass_ = 0 if dw == w else 1
    ratio = dh / h if pass_ == 0 else dw / w
    blur = max(1.0 / ratio, 1.0)
    f = T.t0().filter_generate(name, blur=blur, polar=False)
    fmt = fmt or {1: "r16hf", 2: "rg16hf", 4: "rgba16hf"}[comps]
    got = gpu.simple("ortho", src, fmt, dw, dh, comps=comps, filt=name, sargs_extra=[antiring])
    want = oracle.sample_ortho(src.astype(np.float32), dw, dh, f, pass_, comp_mask=(1 << comps) - 1,
                               coord_mode=T.COORD_EXACT,
                               antiring=None if antiring < 0 else antiring, ratio=ratio)
    return got, want


@pytest.mark.parametrize("name", ORTHO_FILTERS)
@pytest.mark.parametrize("direction", ["vertical", "horizontal"])
def test_ortho_upscale_2x(gpu, oracle, name, direction):
    src = rand_tex(45, 80, 4, seed=31)
    dw, dh = (80, 90) if direction == "vertical" else (160, 45)
    got, want = _ortho_case(gpu, oracle, name, src, dw, dh, 4)
    assert max_ulp32(got, want) <= 1, (name, np.mean(got != want))
    assert np.mean(got != want) < 1e-3


@pytest.mark.parametrize("name", ["lanczos", "mitchell", "hermite"])
@pytest.mark.parametrize("size", [(80, 119), (80, 31), (211, 45), (57, 45)])
def test_ortho_arbitrary_ratio(gpu, oracle, name, size):
    """Non-integer up- and down-scaling (downscaling widens the kernel: blur = 1/ratio)."""
    src = rand_tex(45, 80, 2, seed=32)
    got, want = _ortho_case(gpu, oracle, name, src, size[0], size[1], 2)
    d = np.abs(got - want)
    assert d.max() < 2e-6, (name, size, d.max())


def test_ortho_antiringing(gpu, oracle):
    src = rand_tex(45, 80, 4, seed=33)
    got, want = _ortho_case(gpu, oracle, "lanczos", src, 160, 45, 4, antiring=0.8)
    assert max_ulp32(got, want) <= 1


# ---- built-in fast scalers (sampling.c:319-478) -------------------------------------------
@pytest.mark.parametrize("kind", ["bicubic", "hermite", "gaussian", "oversample"])
@pytest.mark.parametrize("size", [(160, 90), (211, 119), (57, 31)])
def test_builtin_fast_scalers(gpu, oracle, kind, size):
    """pl_shader_sample_bicubic / hermite / gaussian / oversample: up- and down-scaling of an
    80x45 texture, whole and cropped source rects; same fp32 operation sequence on both sides,
    so the comparison is exact (gaussian uses the shared exp definition)."""
    src = rand_tex(45, 80, 4, seed=41)
    extra = [0.2] if kind == "oversample" else []
    for rect in ((0, 0, 0, 0), (3.25, 2.5, 70.0, 40.75)):
        got = gpu.simple(kind, src, "rgba32f", size[0], size[1], comps=4, rect=rect, scale=1.5,
                         sargs_extra=extra)
        want = oracle.sample_fast(src.astype(np.float32), size[0], size[1], kind,
                                  rect=None if not any(rect) else rect, scale=1.5,
                                  threshold=0.2 if kind == "oversample" else 0.0,
                                  coord_mode=T.COORD_EXACT)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), \
            (kind, size, rect, np.abs(got - want).max(), np.mean(got != want))


def test_oversample_without_threshold_and_unorm_source(gpu, oracle):
    src = (rand_tex(30, 40, 2, seed=42) * 65535).astype(np.uint16)
    got = gpu.simple("oversample", src, "rg16", 100, 75, comps=2, sargs_extra=[0.0])
    want = oracle.sample_fast(T.unorm16_to_float(src), 100, 75, "oversample", coord_mode=T.COORD_EXACT)
    assert np.array_equal(got[..., :2].view(np.uint32), want[..., :2].view(np.uint32))


# ---- sampler address modes (pl_sample_src.address_mode) -------------------------------------
@pytest.mark.parametrize("address", [1, 2])  # PL_TEX_ADDRESS_REPEAT, PL_TEX_ADDRESS_MIRROR
@pytest.mark.parametrize("kind", ["bilinear", "nearest", "bicubic"])
def test_address_modes_direct(gpu, oracle, kind, address):
    """A source rect reaching far outside the texture: fetches wrap / mirror per texel index."""
    src = rand_tex(20, 30, 4, seed=51)
    rect = (-41.5, -27.25, 65.0, 52.5)
    got = gpu.simple(kind, src, "rgba32f", 96, 64, comps=4, rect=rect, address=address)
    if kind == "bicubic":
        want = oracle.sample_fast(src.astype(np.float32), 96, 64, kind, rect=rect,
                                  coord_mode=T.COORD_EXACT, address=address)
    else:
        want = oracle.sample_direct(src.astype(np.float32), 96, 64, linear=kind == "bilinear",
                                    rect=rect, coord_mode=T.COORD_EXACT, address=address)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), np.mean(got != want)
    clamped = oracle.sample_direct(src.astype(np.float32), 96, 64, linear=True, rect=rect,
                                   coord_mode=T.COORD_EXACT)
    assert np.mean(clamped != want) > 0.3  # the mode matters for this rect


def test_address_mode_mirror_separable_downscale(gpu, oracle):
    """The contrast-recovery low-pass of the reference (renderer.c:2126-2140): bicubic
    downscaling with a widened kernel and PL_TEX_ADDRESS_MIRROR."""
    src = rand_tex(45, 80, 1, seed=52)
    f = T.t0().filter_generate("bicubic", blur=45 / 13, polar=False)
    got = gpu.simple("ortho", src, "r32f", 80, 13, comps=1, filt="bicubic", sargs_extra=[-1.0], address=2)
    want = oracle.sample_ortho(src.astype(np.float32), 80, 13, f, 0, comp_mask=1,
                               coord_mode=T.COORD_EXACT, ratio=13 / 45, address=2)
    d = np.abs(got[..., 0] - want[..., 0])
    assert d.max() < 2e-6, d.max()
