"""Custom LUTs on the CPU: the .cube parser of the product (pl_lut_parse_cube, csrc/shaders.c)
against the compiled reference's (shaders/lut.c:35-194), call by call on the same text; and
known answers of the oracle's tetrahedral interpolation (the checker of tests/test_gpu_lut.py)."""
import ctypes as C

import numpy as np
import pytest

import plbtest as T


def cube_text(n, dims, seed, domain=None, title=True, crlf=False):
    rng = np.random.default_rng(seed)
    lines = []
    if title:
        lines.append('TITLE "synthetic %d"' % seed)
    lines.append("# a comment line")
    lines.append(("LUT_3D_SIZE %d" if dims == 3 else "LUT_1D_SIZE %d") % n)
    if domain:
        lines.append("DOMAIN_MIN %s %s %s" % tuple(repr(float(v)) for v in domain[0]))
        lines.append("DOMAIN_MAX %s %s %s" % tuple(repr(float(v)) for v in domain[1]))
    lines.append("")
    cnt = n ** 3 if dims == 3 else n
    vals = rng.uniform(-0.25, 1.5, (cnt, 3))
    fmts = ["%.6f", "%.9g", "%e"]
    for i, v in enumerate(vals):
        f = fmts[i % 3]
        lines.append(" ".join(f % x for x in v))
    return ("\r\n" if crlf else "\n").join(lines) + "\n"


def parse(lib, prefix, text):
    b = text.encode()
    size = (C.c_int * 3)()
    cap = 70 * 70 * 70
    data = np.zeros(cap * 3, np.float32)
    fn = getattr(lib, prefix + "lut_parse_cube")
    fn.restype = C.c_int
    n = fn(b, C.c_size_t(len(b)), size, data.ctypes.data_as(C.c_void_p), cap)
    return n, tuple(size), data[:max(n, 0) * 3].copy()


@pytest.mark.parametrize("n,dims,domain,crlf", [
    (17, 3, None, False), (33, 3, ((0, 0, 0), (1, 1, 1)), True), (9, 3, ((-0.5, 0.0, 0.1), (2.0, 1.5, 1.1)), False),
    (64, 1, None, False), (1024, 1, ((0.0, -1.0, 0.0), (4.0, 1.0, 0.5)), True), (2, 3, None, False),
])
def test_cube_parser_matches_reference(n, dims, domain, crlf):
    if not T.have_ref():
        pytest.skip("reference not built")
    ref = C.CDLL(T.REF_SO)
    text = cube_text(n, dims, seed=n + dims, domain=domain, crlf=crlf)
    nr, sr, dr = parse(ref, "hvt_", text)
    nm, sm, dm = parse(T.shim(), "plt_", text)
    assert nr == nm == (n ** 3 if dims == 3 else n)
    assert sr == sm == ((n, n, n) if dims == 3 else (n, 0, 0))
    assert np.array_equal(dr.view(np.uint32), dm.view(np.uint32))


@pytest.mark.parametrize("text", [
    "LUT_3D_SIZE 0\n0 0 0\n", "LUT_3D_SIZE 2\n0 0 0\n1 1 1\n", "DOMAIN_MIN 0 0 0\n0 0 0\n",
    "LUT_1D_SIZE 2\nDOMAIN_MIN 0 0 0\nDOMAIN_MAX 0 1 1\n0 0 0\n1 1 1\n", "LUT_1D_SIZE 2\n0 0 x\n1 1 1\n",
])
def test_cube_parser_rejects_what_the_reference_rejects(text):
    if not T.have_ref():
        pytest.skip("reference not built")
    ref = C.CDLL(T.REF_SO)
    assert parse(ref, "hvt_", text)[0] < 0
    assert parse(T.shim(), "plt_", text)[0] < 0


def identity_lut(n):
    g = np.linspace(0.0, 1.0, n, dtype=np.float32)
    b, gg, r = np.meshgrid(g, g, g, indexing="ij")
    return np.stack([r, gg, b], -1).reshape(-1, 3)


def test_oracle_tetrahedral_known_answers():
    orc = T.oracle()
    rng = np.random.default_rng(5)
    px = rng.uniform(-0.2, 1.2, (64, 64, 4)).astype(np.float32)
    # identity table: the interpolation reproduces the clamped input (to rounding)
    out = orc.custom_lut(px, identity_lut(17), (17, 17, 17))
    assert np.abs(out[..., :3] - np.clip(px[..., :3], 0, 1)).max() < 2e-6
    assert np.array_equal(out[..., 3], px[..., 3])
    # at lattice points the table entry itself comes back, bit for bit
    n = 5
    tab = rng.uniform(0, 1, (n ** 3, 3)).astype(np.float32)
    idx = rng.integers(0, n, (100, 3))
    pts = np.zeros((10, 10, 4), np.float32)
    pts[..., :3] = (idx / np.float32(n - 1)).reshape(10, 10, 3).astype(np.float32)
    got = orc.custom_lut(pts, tab, (n, n, n))[..., :3].reshape(-1, 3)
    want = tab[(idx[:, 2] * n + idx[:, 1]) * n + idx[:, 0]]
    assert np.abs(got - want).max() < 1e-6
    # 1-D: channel k through component k
    t1 = rng.uniform(0, 1, (32, 3)).astype(np.float32)
    x = np.zeros((1, 32, 4), np.float32)
    x[0, :, :3] = (np.arange(32) / np.float32(31))[:, None]
    got = orc.custom_lut(x, t1, (32, 0, 0))[0, :, :3]
    assert np.abs(got - t1).max() < 1e-6
