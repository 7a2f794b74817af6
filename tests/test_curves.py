"""Transfer curves (csrc/curves.h, oracle/pl_oracle_curve.h) and the fp32 hot-path functions.

CPU only. Three independent statements have to agree:
  * the product's table generator (curves.c, through pl_cuda_curve_table),
  * the oracle's restatement (orc_curve_table)                         -> bit for bit,
  * a float64 closed form of the reference's formulas (numpy)          -> within the stated bound.
The bound is what makes the redefinition legitimate: every curve is within 3 float ULP of the real
function (1.6 ULP inside the nominal signal range), i.e. far below 1/4 LSB of a 16-bit output, and
closer to the real function than an fp32 pow/divide/pow chain.
"""
import ctypes as C

import numpy as np
import pytest

import plbtest as T

f32 = np.float32
CURVES = ["PQ_EOTF_SDR", "PQ_EOTF", "PQ_OETF_SDR", "PQ_OETF_DELIN", "PQ_OETF", "PQX_OETF",
          "POW_24", "POW_INV24", "POW_22", "POW_INV22"]


def glsl6(x):
    return float("%f" % x)


def _consts():
    c = {}
    m1, m2 = float(f32(2610. / 4096 / 4)), float(f32(2523. / 4096 * 128))
    c1, c2, c3 = float(f32(3424. / 4096)), float(f32(2413. / 4096 * 32)), float(f32(2392. / 4096 * 32))
    c["exact"] = (m1, m2, c1, c2, c3, 1 / m1, 1 / m2)
    c["lit"] = (float(f32(glsl6(m1))), float(f32(glsl6(m2))), float(f32(glsl6(c1))),
                float(f32(glsl6(c2))), float(f32(glsl6(c3))), float(f32(1 / glsl6(m1))),
                float(f32(1 / glsl6(m2))))
    c["to_pq"] = float(f32(glsl6(203 / 10000)))
    c["from_pq"] = float(f32(glsl6(10000 / 203)))
    c["inv_from_pq"] = float(f32(1 / glsl6(10000 / 203)))
    return c


def real_function(i, x):
    """float64 closed form of curve i (src/shaders/colorspace.c:662-668, :778-784, :622-630)."""
    k = _consts()

    def eotf(x, s, q=k["lit"]):
        m1, m2, c1, c2, c3, im1, im2 = q
        p = x ** im2
        return s * (np.maximum(p - c1, 0) / (c2 - c3 * p)) ** im1

    def oetf(x, s, q=k["lit"]):
        m1, m2, c1, c2, c3, im1, im2 = q
        u = (s * x) ** m1
        return ((c1 + c2 * u) / (1 + c3 * u)) ** m2

    return [lambda: eotf(x, k["from_pq"]), lambda: eotf(x, 1.0), lambda: oetf(x, k["to_pq"]),
            lambda: oetf(x, k["inv_from_pq"]), lambda: oetf(x, 1.0), lambda: oetf(x, 1.0, k["exact"]),
            lambda: x ** float(f32(2.4)), lambda: x ** float(f32(1 / 2.4)),
            lambda: x ** float(f32(2.2)), lambda: x ** float(f32(1 / 2.2))][i]()


def product_table(i):
    lib = C.CDLL(T.PRODUCT_SO)
    lib.pl_cuda_curve_table.restype = C.c_bool
    b, n, m = C.c_int(), C.c_int(), C.c_int()
    assert lib.pl_cuda_curve_table(i, C.byref(b), C.byref(n), C.byref(m), None)
    coef = np.zeros((n.value + 1, 4), f32)
    assert lib.pl_cuda_curve_table(i, C.byref(b), C.byref(n), C.byref(m), T.fp(coef))
    return b.value, n.value, m.value, coef


def oracle_table(i):
    lib = T.oracle().lib
    b, n, m = C.c_int(), C.c_int(), C.c_int()
    assert lib.orc_curve_table(i, C.byref(b), C.byref(n), C.byref(m), None)
    coef = np.zeros((n.value + 1, 4), f32)
    assert lib.orc_curve_table(i, C.byref(b), C.byref(n), C.byref(m), T.fp(coef))
    return b.value, n.value, m.value, coef


def evaluate(tab, x):
    """numpy restatement of plm_curve / orc_curve_eval (fp32 Horner with fused multiply-adds)."""
    base, nseg, m, coef = tab
    ix = x.view(np.uint32)
    idx = np.maximum((ix >> np.uint32(23 - m)).astype(np.int64), base) - base
    hit = idx <= nseg
    k = coef[np.where(hit, idx, 0)]
    u = (((ix << np.uint32(m)) & np.uint32(0x7FFFFF)) | np.uint32(0x3F800000)).view(f32) - f32(1.5)

    def fma(a, b, c):
        return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)

    return fma(fma(fma(k[:, 3], u, k[:, 2]), u, k[:, 1]), u, k[:, 0]), hit


@pytest.mark.parametrize("i", range(len(CURVES)), ids=CURVES)
def test_product_and_oracle_tables_are_identical(i):
    pb, pn, pm, pc = product_table(i)
    ob, on, om, oc = oracle_table(i)
    assert (pb, pn, pm) == (ob, on, om)
    assert np.array_equal(pc.view(np.uint32), oc.view(np.uint32))


@pytest.mark.parametrize("i", range(len(CURVES)), ids=CURVES)
def test_curve_is_within_bound_of_the_real_function(i):
    tab = product_table(i)
    base, nseg, m, _ = tab
    lo, hi = (base + 1) << (23 - m), (base + 1 + nseg) << (23 - m)
    rng = np.random.default_rng(i)
    bits = rng.integers(lo, hi, 2_000_000, dtype=np.uint32)
    # every segment boundary and its neighbours
    edges = (np.arange(base + 1, base + 1 + nseg, dtype=np.uint32) << np.uint32(23 - m))
    bits = np.concatenate([bits, edges, edges[1:] - 1, edges + 1])
    x = bits.view(f32)
    y, hit = evaluate(tab, x)
    assert hit.all()
    want = real_function(i, x.astype(np.float64))
    # absolute floor: outputs below 1e-13 of the 10 000-nit scale are 1e-9 nits
    err = np.abs(y - want)
    ulp = np.spacing(np.maximum(np.abs(want), 1e-13).astype(f32)).astype(np.float64)
    worst = (err / ulp).max()
    nominal = x <= 1.0
    worst_nominal = (err[nominal] / ulp[nominal]).max()
    assert worst <= 3.0, f"{CURVES[i]}: {worst:.2f} ULP"
    assert worst_nominal <= 1.75, f"{CURVES[i]}: {worst_nominal:.2f} ULP inside [0, 1]"


@pytest.mark.parametrize("i", range(len(CURVES)), ids=CURVES)
def test_oracle_curve_apply_matches_table_and_falls_back_outside(i):
    """orc_curve_apply == table inside the domain, == the exact pow chain outside (and everywhere
    in exact mode)."""
    o = T.oracle()
    tab = oracle_table(i)
    base, nseg, m, _ = tab
    hi = f32(np.array([(base + 1 + nseg) << (23 - m)], np.uint32).view(f32)[0])
    rng = np.random.default_rng(100 + i)
    x = np.concatenate([rng.uniform(0, float(hi) * 1.3, 20000), [0.0, 1e-30, float(hi)]]).astype(f32)
    got = np.empty_like(x)
    o.lib.orc_curve_apply(i, T.fp(x), T.fp(got), C.c_size_t(x.size))
    y, hit = evaluate(tab, x)
    assert np.array_equal(got[hit].view(np.uint32), y[hit].view(np.uint32))
    assert (~hit).any()
    o.lib.orc_set_exact_math(1)
    try:
        exact = np.empty_like(x)
        o.lib.orc_curve_apply(i, T.fp(x), T.fp(exact), C.c_size_t(x.size))
    finally:
        o.lib.orc_set_exact_math(0)
    assert np.array_equal(got[~hit].view(np.uint32), exact[~hit].view(np.uint32))
    # curve vs the round-1 chain of correctly rounded fp32 pow(): the OETF and the powers agree to
    # a few ULP; the EOTF chain loses up to ~2e-4 (relative) to the cancellation in
    # (x^(1/m2) - c1), which the curve — fitted to the real function — does not have
    lo = np.array([(base + 1) << (23 - m)], np.uint32).view(f32)[0]
    ok = hit & (exact > 1e-12) & (x >= lo)
    rel = np.abs(got[ok].astype(np.float64) - exact[ok]) / exact[ok]
    assert rel.max() < (5e-4 if i < 2 else 3e-5 if i < 6 else 2e-6)
    # below the domain the curve is the constant f(0): absolute error < 1e-8
    below = hit & (x < lo)
    assert np.abs(got[below].astype(np.float64) - exact[below]).max() < 1e-8


def test_fast_atan2_and_sincos_accuracy():
    o = T.oracle()
    rng = np.random.default_rng(7)
    n = 1 << 20
    a = rng.normal(0, 1, n).astype(f32) * f32(0.3)
    b = rng.normal(0, 1, n).astype(f32) * f32(0.3)
    a[:8] = [0, 0, 1, -1, 1e-30, -0.0, np.inf, np.nan]
    b[:8] = [0, -1, 0, 0, 1e-30, -0.0, 1, 1]
    out = np.empty(n, f32)
    o.lib.orc_math(8, T.fp(a), T.fp(b), T.fp(out), C.c_size_t(n))
    want = np.arctan2(a.astype(np.float64), b.astype(np.float64))
    fin = np.isfinite(want)
    assert np.abs(out[fin] - want[fin]).max() < 3.5e-7      # ~1.5 ULP at pi
    assert np.isnan(out[7])
    x = rng.uniform(-64, 64, n).astype(f32)
    x[:3] = [0, 6.283185, -6.283185]
    for fn, ref in ((6, np.sin), (7, np.cos)):
        o.lib.orc_math(fn, T.fp(x), T.fp(x), T.fp(out), C.c_size_t(n))
        # argument reduction with a two-term pi/2: error grows with |x|; deband uses [0, 2 pi)
        small = np.abs(x) <= 6.3
        assert np.abs(out[small] - ref(x[small].astype(np.float64))).max() < 1.5e-7
        assert np.abs(out - ref(x.astype(np.float64))).max() < 3e-6
