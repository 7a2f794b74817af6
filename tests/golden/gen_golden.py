#!/usr/bin/env python
"""Generates tests/golden/*.npz from the COMPILED REFERENCE (oracle/_ref, built from
/root/reference by oracle/build_ref.sh). Run here (the reference does not exist on the GPU
box); the fixtures travel with the repository and pin the product's host tables and the
oracle's inputs without the reference being present.

    python tests/golden/gen_golden.py

Contents
  filters.npz     pl_filter_generate weights / geometry for the scalers the path uses
  tone.npz        pl_tone_map_generate LUTs (256 entries) for every tone-mapping function,
                  HDR10 1000 nit -> SDR 203 nit, plus the spline curve for a measured peak
  gamut.npz       every 8th hue plane of the pl_gamut_map_generate 3DLUT (48x32x256), perceptual BT.2020 -> BT.709
  dither.npz      blue noise 64x64 (fresh-process result) and Bayer 64x64
  color.npz       pl_color_repr_decode matrices, raw primaries, RGB<->XYZ / LMS matrices
  h274_db.npz     the 832x832 H.274 grain data base (as int8 = value * 255)
  av1_grain.npz   LUTs uploaded by the real pl_shader_film_grain for the reference's fixture
"""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import plbtest as T  # noqa: E402


def main():
    ref = T.ref()
    assert ref is not None, "build the reference first (oracle/build_ref.sh)"
    out = {}
    # ---- filters
    f = {}
    for name, polar in (("ewa_lanczos", True), ("ewa_lanczossharp", True), ("lanczos", False),
                        ("mitchell", False), ("hermite", False), ("gaussian", False)):
        g = ref.filter_generate(name, polar=polar)
        f[name + "_weights"] = g["weights"]
        f[name + "_geom"] = np.array([g["radius"], g["radius_zero"], g.get("row_size", 0),
                                      g.get("row_stride", 0)], np.float64)
    np.savez_compressed(os.path.join(HERE, "filters.npz"), **f)
    # ---- tone curves
    hdr = T.csp(T.PRIM["BT_2020"], T.TRC["PQ"], max_luma=1000)
    sdr = T.csp(T.PRIM["BT_709"], T.TRC["BT_1886"])
    imin, imax, iavg = ref.nominal_luma(ref.csp_infer(hdr), T.META["ANY"], T.HDR["PQ"])
    omin, omax, _ = ref.nominal_luma(ref.csp_infer(sdr), T.META["HDR10"], T.HDR["PQ"])
    t = {"range": np.array([imin, imax, iavg, omin, omax], np.float32)}
    for fn in T.TONE_FNS:
        t[fn] = ref.tone_map_generate(fn, imin, imax, iavg, omin, omax, hdr_csp=hdr)["lut"]
    measured = hdr.copy()
    measured[19], measured[18] = 0.7688, 0.99999
    t["spline_measured"] = ref.tone_map_generate("spline", imin, np.float32(0.99999), np.float32(0.7688),
                                                 omin, omax, hdr_csp=measured)["lut"]
    np.savez_compressed(os.path.join(HERE, "tone.npz"), **t)
    # ---- gamut
    gmin, gmax, _ = ref.nominal_luma(ref.csp_infer(sdr), T.META["HDR10"], T.HDR["PQ"])
    src8, dst8 = ref.raw_primaries(T.PRIM["BT_2020"]), ref.raw_primaries(T.PRIM["BT_709"])
    dst8c, _ = ref.primaries_clip(dst8, src8)
    lut = ref.gamut_map_generate("perceptual", src8, dst8c, gmin, gmax)["lut"]
    # every 8th hue plane keeps the fixture small (the full LUT is 48x32x256x3 floats)
    np.savez_compressed(os.path.join(HERE, "gamut.npz"),
                        perceptual_2020_709_h8=lut.reshape(256, 32, 48, 3)[::8].astype(np.float32),
                        luma=np.array([gmin, gmax], np.float32))
    # ---- dither (blue noise needs a fresh process: the generator consumes libc rand())
    code = ("import sys; sys.path.insert(0, %r); import plbtest as T; "
            "sys.stdout.buffer.write(T.ref().dither_matrix(0, 64, fresh=True).tobytes())" % os.path.dirname(HERE))
    bn = np.frombuffer(subprocess.run([sys.executable, "-c", code], capture_output=True, check=True).stdout,
                       np.float32).reshape(64, 64)
    np.savez_compressed(os.path.join(HERE, "dither.npz"), blue_noise_64=bn,
                        bayer_64=ref.dither_matrix(1, 64, fresh=True))
    # ---- colour
    c = {}
    for sysn, lv, sd, cd in (("BT_709", "LIMITED", 16, 10), ("BT_2020_NC", "LIMITED", 16, 10),
                              ("BT_601", "FULL", 8, 8), ("YCGCO", "FULL", 8, 8)):
        d = ref.repr_decode(T.SYS[sysn], T.LEVELS[lv], sd, cd)
        c[f"decode_{sysn}_{lv}_{sd}_{cd}"] = np.concatenate([d["m"], d["c"], [d["scale"]]]).astype(np.float32)
    for p in ("BT_709", "BT_2020", "DISPLAY_P3"):
        raw = ref.raw_primaries(T.PRIM[p])
        c["prim_" + p] = raw
        for which in range(4):
            c[f"matrix{which}_{p}"] = ref.matrix(which, raw)
    np.savez_compressed(os.path.join(HERE, "color.npz"), **c)
    # ---- AV1 grain LUTs harvested from the real shader
    hv = C.CDLL(os.path.join(T.ROOT, "oracle", "_ref", "libplref_harvest.so"))
    g = {}
    for plane in (0, 1):
        d = tempfile.mkdtemp()
        assert hv.hvs_av1_grain(d.encode(), 256, 144, plane, 1234, 1, 10, 1, 2) == 0
        for line in open(os.path.join(d, "meta.txt")):
            fl = line.split()
            if fl and fl[0] == "desc" and fl[-1] == "1":
                dt = np.uint32 if fl[4] == "r32u" else np.float32
                g[f"plane{plane}_desc{fl[1]}"] = np.fromfile(os.path.join(d, f"desc_{fl[1]}.bin"), dt)
    np.savez_compressed(os.path.join(HERE, "av1_grain.npz"), **g)
    # ---- H.274 grain data base filled by the real sh_lut callback (values are n / 255)
    sys.path.insert(0, os.path.join(T.ROOT, "tests"))
    from test_grain_tables import harvest_h274
    db, _ = harvest_h274()
    q = np.rint(db.astype(np.float64) * 255).astype(np.int8)
    assert np.array_equal((q.astype(np.float64) / 255.0).astype(np.float32), db)
    np.savez_compressed(os.path.join(HERE, "h274_db.npz"), db_times_255=q)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
