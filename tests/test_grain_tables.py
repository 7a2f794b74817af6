"""AV1 film grain host tables (libplacebo_b200/csrc/film_grain.c) against the LUTs the real
pl_shader_film_grain uploads (harvested from the compiled reference on its dummy GPU), for the
reference's own fixture (src/tests/utils.h:240-262). Integer algorithm -> bit-exact."""
import ctypes as C
import os
import tempfile

import numpy as np
import pytest

import plbtest as T

HARVEST = os.path.join(T.ROOT, "oracle", "_ref", "libplref_harvest.so")
pytestmark = pytest.mark.skipif(not os.path.exists(HARVEST), reason="reference harvester not built")


def harvest(plane, w, h, seed, depth):
    hv = C.CDLL(HARVEST)
    d = tempfile.mkdtemp()
    assert hv.hvs_av1_grain(d.encode(), w, h, plane, seed, 1, depth, 1, 2) == 0
    descs = {}
    for line in open(os.path.join(d, "meta.txt")):
        f = line.split()
        if f and f[0] == "desc" and f[-1] == "1":
            descs[int(f[1])] = (f[4], int(f[5]), int(f[6]), os.path.join(d, f"desc_{f[1]}.bin"))
    return descs


def product_tables(seed, depth, sub, w, h):
    sh = T.shim()
    gy, gcb, gcr = np.zeros(4096, np.float32), np.zeros(4096 >> (2 * sub), np.float32), np.zeros(4096 >> (2 * sub), np.float32)
    ow, oh = ((w << sub) + 127) // 128 * 4, ((h << sub) + 127) // 128 * 4
    off = np.zeros(ow * oh, np.uint32)
    sc = np.zeros(3 * 256, np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    assert sh.shim_av1_tables(seed, depth, 1, 2, sub, sub, w, h, p(gy), p(gcb), p(gcr), p(off), p(sc)) == 0
    return gy, gcb, gcr, off.reshape(oh, ow), sc.reshape(3, 256)


@pytest.mark.parametrize("seed,depth", [(1234, 10), (7, 8), (65535, 12)])
def test_av1_tables_match_reference(seed, depth):
    w, h = 256, 144
    gy, _, _, off, sc = product_tables(seed, depth, 0, w, h)
    d = harvest(0, w, h, seed, depth)
    ref_grain = np.fromfile(d[0][3], np.float32)
    ref_off = np.fromfile(d[1][3], np.uint32).reshape(d[1][2], d[1][1])
    ref_sc = np.fromfile(d[2][3], np.float32)
    assert np.array_equal(gy, ref_grain)
    assert np.array_equal(off, ref_off)
    assert np.array_equal(sc[0], ref_sc)

    _, gcb, gcr, off_c, sc_c = product_tables(seed, depth, 1, w // 2, h // 2)
    d = harvest(1, w, h, seed, depth)
    ref_uv = np.fromfile(d[0][3], np.float32).reshape(-1, 2)
    assert np.array_equal(gcb, ref_uv[:, 0]) and np.array_equal(gcr, ref_uv[:, 1])
    assert np.array_equal(off_c, np.fromfile(d[1][3], np.uint32).reshape(d[1][2], d[1][1]))
    assert np.array_equal(sc_c[1], np.fromfile(d[2][3], np.float32))
    assert np.array_equal(sc_c[2], np.fromfile(d[3][3], np.float32))


# ---- H.274 / SMPTE RDD 5 grain data base (film_grain_h274.c:33-125) ------------------------
def harvest_h274(fmt=b"r16", w=64, h=64, comps=1, cmap=(0, 0, 0, 0), seed=1234, depth=10, levels=1, sys_=2):
    """(data base 832x832 float32, {uniform name: values}) of the real pl_shader_film_grain
    (H.274) for the reference's fixture (src/tests/utils.h:263-277)."""
    hv = C.CDLL(HARVEST)
    d = tempfile.mkdtemp()
    m = (C.c_int * 4)(*cmap)
    rc = hv.hvs_h274_grain(d.encode(), fmt, w, h, comps, m, seed, depth, levels, sys_)
    if rc == -3:
        return None, None
    assert rc == 0
    uni = []
    for line in open(os.path.join(d, "meta.txt")):
        f = line.split()
        if f and f[0] == "var":
            uni.append([float(x) for x in f[f.index(":") + 1:]])
    return np.fromfile(os.path.join(d, "desc_0.bin"), np.float32).reshape(832, 832), uni


def product_h274_db():
    db = np.zeros((832, 832), np.float32)
    assert T.shim().shim_h274_database(db.ctypes.data_as(C.c_void_p)) == 0
    return db


def test_h274_database_matches_reference():
    ref, uni = harvest_h274()
    got = product_h274_db()
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    # the uniforms of the fixture: scale_factor 2^-8, seed, bounds / 255, scale<<16 | h<<8 | v
    assert uni[0] == [2.0 ** -8] and uni[1] == [1234.0]
    assert np.allclose(uni[2], [10 / 255, 250 / 255], rtol=1e-7) and uni[3] == [16 << 16 | 10 << 8 | 12]
