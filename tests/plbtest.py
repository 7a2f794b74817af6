"""Shared test plumbing: ctypes bindings of the CPU oracle (oracle/libpl_oracle.so), of the
reference harvester (oracle/_ref/libplref_harvest.so, optional) and of the C test driver
(tests/c/libplb_shim.so) that exercises the public C ABI of libplacebo_b200.so."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

f32 = np.float32
FP = C.POINTER(C.c_float)


def fp(a):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(FP)


def ip(a):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_int))


# ----------------------------------------------------------------------------------------
# enums (include/libplacebo_b200.h)
# ----------------------------------------------------------------------------------------
SYS = dict(UNKNOWN=0, BT_601=1, BT_709=2, SMPTE_240M=3, BT_2020_NC=4, BT_2020_C=5, BT_2100_PQ=6,
           BT_2100_HLG=7, DOLBYVISION=8, YCGCO=9, YCGCO_RE=10, YCGCO_RO=11, RGB=12, XYZ=13)
LEVELS = dict(UNKNOWN=0, LIMITED=1, FULL=2)
PRIM = dict(UNKNOWN=0, BT_601_525=1, BT_601_625=2, BT_709=3, BT_470M=4, EBU_3213=5, BT_2020=6,
            APPLE=7, ADOBE=8, PRO_PHOTO=9, CIE_1931=10, DCI_P3=11, DISPLAY_P3=12, V_GAMUT=13,
            S_GAMUT=14, FILM_C=15, ACES_AP0=16, ACES_AP1=17)
TRC = dict(UNKNOWN=0, BT_1886=1, SRGB=2, LINEAR=3, GAMMA18=4, GAMMA20=5, GAMMA22=6, GAMMA24=7,
           GAMMA26=8, GAMMA28=9, PRO_PHOTO=10, ST428=11, PQ=12, HLG=13, V_LOG=14, S_LOG1=15,
           S_LOG2=16, SCRGB=17)
HDR = dict(NORM=0, SQRT=1, NITS=2, PQ=3)
META = dict(ANY=0, NONE=1, HDR10=2, HDR10PLUS=3, CIE_Y=4)
TONE_FNS = ["clip", "st2094-40", "st2094-10", "bt2390", "bt2446a", "spline", "reinhard",
            "mobius", "hable", "gamma", "linear", "linearlight"]
ED_KERNELS = ["simple", "false-fs", "sierra-lite", "floyd-steinberg", "atkinson",
              "jarvis-judice-ninke", "stucki", "burkes", "sierra-2", "sierra-3"]
GAMUT_FNS = ["clip", "perceptual", "softclip", "relative", "saturation", "absolute",
             "desaturate", "darken", "highlight", "linear"]
TONE_CONSTANTS = np.array([0.4, 0.1, 0.8, 0.4, 1.0, 1.5, 0.2, 0.5, 0.5, 0.3, 1.0], f32)
# struct order: perceptual_deadzone, perceptual_strength, colorimetric_gamma, knee, desat
GAMUT_CONSTANTS = np.array([0.30, 0.80, 1.80, 0.70, 0.35], f32)
COORD_LITERAL, COORD_EXACT = 0, 1


def csp(prim=0, trc=0, min_luma=0.0, max_luma=0.0, max_cll=0.0, max_fall=0.0,
        scene_max=(0, 0, 0), scene_avg=0.0, max_pq_y=0.0, avg_pq_y=0.0, hdr_prim=None):
    """Flat pl_color_space (layout documented in oracle/t0_flat.inc)."""
    a = np.zeros(20, f32)
    a[0], a[1] = prim, trc
    if hdr_prim is not None:
        a[2:10] = hdr_prim
    a[10], a[11], a[12], a[13] = min_luma, max_luma, max_cll, max_fall
    a[14:17] = scene_max
    a[17], a[18], a[19] = scene_avg, max_pq_y, avg_pq_y
    return a


# ----------------------------------------------------------------------------------------
# building / loading
# ----------------------------------------------------------------------------------------
ORACLE_SO = os.path.join(ROOT, "oracle", "libpl_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libplref_harvest.so")
SHIM_SO = os.path.join(ROOT, "tests", "c", "libplb_shim.so")
PRODUCT_SO = os.path.join(ROOT, "libplacebo_b200", "libplacebo_b200.so")

_cache = {}


def _ensure_built():
    if _cache.get("built"):
        return
    need = [ORACLE_SO, SHIM_SO, PRODUCT_SO]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__
        __graft_entry__.build()
    _cache["built"] = True


def oracle():
    _ensure_built()
    if "oracle" not in _cache:
        _cache["oracle"] = Oracle(C.CDLL(ORACLE_SO))
    return _cache["oracle"]


def have_ref():
    return os.path.exists(REF_SO)


def ref():
    """Flat T0 API of the real reference (hvt_*), or None if oracle/_ref was never built."""
    if "ref" not in _cache:
        _cache["ref"] = T0(C.CDLL(REF_SO), "hvt_") if have_ref() else None
    return _cache["ref"]


def product_t0():
    """Flat T0 API of libplacebo_b200 (plt_*): host code only, no GPU needed."""
    _ensure_built()
    if "pt0" not in _cache:
        import libplacebo_b200
        libplacebo_b200.load()
        _cache["pt0"] = T0(C.CDLL(SHIM_SO), "plt_")
    return _cache["pt0"]


def t0():
    """Parameter source for the oracle: the real reference when present, else the product's
    host code (which tests/test_t0_*.py pin bit-exactly against the reference)."""
    return ref() or product_t0()


def shim():
    _ensure_built()
    if "shim" not in _cache:
        import libplacebo_b200
        libplacebo_b200.load()
        _cache["shim"] = C.CDLL(SHIM_SO)
    return _cache["shim"]


# ----------------------------------------------------------------------------------------
# flat T0 API (oracle/t0_flat.inc) — identical for reference and product
# ----------------------------------------------------------------------------------------
class T0:
    def __init__(self, lib, prefix):
        self.lib, self.prefix = lib, prefix
        self.f("hdr_rescale").restype = C.c_float
        self.f("filter_sample").restype = C.c_double

    def f(self, name):
        return getattr(self.lib, self.prefix + name)

    def filter_generate(self, name, blur=0.0, polar=None, lut_entries=256):
        polar = name.startswith("ewa") if polar is None else polar
        w = np.zeros(lut_entries * 64, f32)
        meta = np.zeros(8, f32)
        n = self.f("filter_generate")(name.encode(), C.c_float(blur), lut_entries,
                                      C.c_float(1e-3 if polar else 0.0), 0 if polar else 1024,
                                      0 if polar else 4, fp(w), w.size, fp(meta))
        if n < 0:
            return None
        return dict(weights=w[:n].copy(), radius=meta[0], radius_zero=meta[1],
                    row_size=int(meta[2]), row_stride=int(meta[3]), polar=bool(meta[4]),
                    antiring=meta[5], cfg_radius=meta[6], blur=meta[7])

    def filter_sample(self, name, x, blur=0.0):
        return self.f("filter_sample")(name.encode(), C.c_double(blur), C.c_double(x))

    def repr_decode(self, sys, levels, sample_depth, color_depth, bit_shift=0):
        o = np.zeros(16, f32)
        self.f("repr_decode")(sys, levels, sample_depth, color_depth, bit_shift, fp(o))
        return dict(m=o[:9].copy(), c=o[9:12].copy(), scale=o[12], raw=o)

    def repr_encode(self, sys, levels, sample_depth, color_depth, bit_shift=0):
        o = np.zeros(16, f32)
        self.f("repr_encode")(sys, levels, sample_depth, color_depth, bit_shift, fp(o))
        return dict(m=o[:9].copy(), c=o[9:12].copy(), scale=o[12], raw=o)

    def repr_decode_dovi(self, sample_depth, color_depth, offset3, nonlinear9, linear9):
        o = np.zeros(22, f32)
        d = np.concatenate([np.asarray(offset3, f32), np.asarray(nonlinear9, f32)]).astype(f32)
        l = np.ascontiguousarray(linear9, f32)
        self.f("repr_decode_dovi")(sample_depth, color_depth, fp(d), fp(l), fp(o))
        return dict(scale=o[0], m=o[1:10].copy(), c=o[10:13].copy(), lms2rgb=o[13:22].copy())

    def raw_primaries(self, prim):
        o = np.zeros(8, f32)
        self.f("raw_primaries")(prim, fp(o))
        return o

    def matrix(self, which, prim8):
        o = np.zeros(9, f32)
        self.f("matrix")(which, fp(np.ascontiguousarray(prim8, f32)), fp(o))
        return o

    def color_mapping_matrix(self, src8, dst8, intent):
        o = np.zeros(9, f32)
        self.f("color_mapping_matrix")(fp(src8), fp(dst8), intent, fp(o))
        return o

    def hdr_rescale(self, frm, to, x):
        return f32(self.f("hdr_rescale")(frm, to, C.c_float(x)))

    def csp_infer(self, c):
        c = c.copy()
        self.f("csp_infer")(fp(c))
        return c

    def csp_infer_map(self, s, d):
        s, d = s.copy(), d.copy()
        self.f("csp_infer_map")(fp(s), fp(d))
        return s, d

    def nominal_luma(self, c, metadata, scaling):
        o = np.zeros(3, f32)
        self.f("nominal_luma")(fp(c), metadata, scaling, fp(o))
        return o

    def tone_map_generate(self, fn, input_min, input_max, input_avg, output_min, output_max,
                          lut_size=256, scaling=(3, 3), constants=TONE_CONSTANTS, hdr_csp=None,
                          infer=True):
        par = np.concatenate([np.array([scaling[0], scaling[1], lut_size, input_min, input_max,
                                        input_avg, output_min, output_max], f32), constants]).astype(f32)
        lut = np.zeros(lut_size, f32)
        h = hdr_csp if hdr_csp is not None else csp()
        noop = self.f("tone_map_generate")(fn.encode(), fp(par), fp(h), int(infer), fp(lut))
        return dict(lut=lut, noop=bool(noop), params=par)

    def gamut_map_generate(self, fn, src8, dst8, min_luma, max_luma, size=(48, 32, 256),
                           constants=GAMUT_CONSTANTS, generate=True):
        par = np.concatenate([src8, dst8, np.array([min_luma, max_luma], f32), constants,
                              np.array(size, f32)]).astype(f32)
        lut = np.zeros(size[0] * size[1] * size[2] * 3, f32) if generate else None
        noop = self.f("gamut_map_generate")(fn.encode(), fp(par), fp(lut) if generate else None)
        return dict(lut=lut, noop=bool(noop))

    def primaries_clip(self, src8, dst8):
        o = np.zeros(8, f32)
        compat = self.f("primaries_clip")(fp(src8), fp(dst8), fp(o))
        return o, bool(compat)

    def primaries_compatible(self, a8, b8):
        return bool(self.f("primaries_compatible")(fp(np.ascontiguousarray(a8, f32)),
                                                   fp(np.ascontiguousarray(b8, f32))))

    _dither_cache = {}

    def dither_matrix(self, method, size, fresh=False):
        """The reference's blue-noise generator breaks ties with libc rand() and never seeds
        it, so only the FIRST call in a process is reproducible (fresh-process behaviour, which
        is what the product reproduces). Cache that first result per library."""
        key = (self.prefix, method, size)
        if fresh or key not in T0._dither_cache:
            o = np.zeros(size * size, f32)
            self.f("dither_matrix")(method, size, fp(o))
            if fresh:
                return o.reshape(size, size)
            T0._dither_cache[key] = o.reshape(size, size)
        return T0._dither_cache[key].copy()

    def error_diffusion_kernel(self, name):
        o = np.zeros(17, np.int32)
        r = self.f("error_diffusion_kernel")(name.encode(), ip(o))
        return None if r else dict(shift=int(o[0]), divisor=int(o[1]), pattern=o[2:].reshape(3, 5).copy())


# ----------------------------------------------------------------------------------------
# oracle bindings
# ----------------------------------------------------------------------------------------
class OrcTex(C.Structure):
    _fields_ = [("w", C.c_int), ("h", C.c_int), ("comps", C.c_int), ("data", FP), ("addr", C.c_int)]


class OrcRect(C.Structure):
    _fields_ = [("x0", C.c_float), ("y0", C.c_float), ("x1", C.c_float), ("y1", C.c_float)]


class OrcPolar(C.Structure):
    _fields_ = [("lut", FP), ("lut_entries", C.c_int), ("radius", C.c_float),
                ("radius_zero", C.c_float), ("antiring", C.c_float), ("scale", C.c_float),
                ("comp_mask", C.c_uint8), ("coord_mode", C.c_int)]


class OrcOrtho(C.Structure):
    _fields_ = [("lut", FP), ("lut_entries", C.c_int), ("row_size", C.c_int),
                ("row_stride", C.c_int), ("use_linear", C.c_int), ("antiring", C.c_float),
                ("ratio", C.c_float), ("scale", C.c_float), ("comp_mask", C.c_uint8),
                ("pass_", C.c_int), ("coord_mode", C.c_int)]


class OrcDeband(C.Structure):
    _fields_ = [("iterations", C.c_int), ("threshold", C.c_float), ("radius", C.c_float),
                ("grain", C.c_float), ("grain_neutral", C.c_float * 3), ("scale", C.c_float),
                ("comp_mask", C.c_uint8), ("frame_index", C.c_uint32)]


class OrcDoviComp(C.Structure):
    _fields_ = [("num_pivots", C.c_int), ("pivots", C.c_float * 9), ("method", C.c_int * 8),
                ("poly_coeffs", (C.c_float * 3) * 8), ("mmr_order", C.c_int * 8),
                ("mmr_constant", C.c_float * 8), ("mmr_coeffs", ((C.c_float * 7) * 3) * 8)]


class DoviComp(C.Structure):  # struct pl_reshape_data of the public header
    _fields_ = [("num_pivots", C.c_uint8), ("pivots", C.c_float * 9), ("method", C.c_uint8 * 8),
                ("poly_coeffs", (C.c_float * 3) * 8), ("mmr_order", C.c_uint8 * 8),
                ("mmr_constant", C.c_float * 8), ("mmr_coeffs", ((C.c_float * 7) * 3) * 8)]


class DoviMeta(C.Structure):  # struct pl_dovi_metadata
    _fields_ = [("nonlinear_offset", C.c_float * 3), ("nonlinear", C.c_float * 9),
                ("linear", C.c_float * 9), ("comp", DoviComp * 3)]


def dovi_fixture(variant=0):
    """The reference's Dolby Vision test fixture (src/tests/utils.h:279-323) as held by the C shim."""
    sh = shim()
    sh.shim_dovi_fixture_ptr.restype = C.c_void_p
    return DoviMeta.from_address(sh.shim_dovi_fixture_ptr(variant))


class Av1GrainData(C.Structure):  # struct pl_av1_grain_data of the public header
    _fields_ = [("num_points_y", C.c_int), ("points_y", (C.c_uint8 * 2) * 14),
                ("chroma_scaling_from_luma", C.c_bool), ("num_points_uv", C.c_int * 2),
                ("points_uv", ((C.c_uint8 * 2) * 10) * 2), ("scaling_shift", C.c_int),
                ("ar_coeff_lag", C.c_int), ("ar_coeffs_y", C.c_int8 * 24),
                ("ar_coeffs_uv", (C.c_int8 * 25) * 2), ("ar_coeff_shift", C.c_int),
                ("grain_scale_shift", C.c_int), ("uv_mult", C.c_int8 * 2),
                ("uv_mult_luma", C.c_int8 * 2), ("uv_offset", C.c_int16 * 2), ("overlap", C.c_bool)]


class OrcAv1Grain(C.Structure):
    _fields_ = [("grain", FP * 3), ("grain_w", C.c_int), ("grain_h", C.c_int),
                ("offsets", C.POINTER(C.c_uint32)), ("offsets_w", C.c_int), ("offsets_h", C.c_int),
                ("scaling", FP * 3), ("is_luma", C.c_int * 3), ("num_comps", C.c_int),
                ("sub_x", C.c_int), ("sub_y", C.c_int), ("overlap", C.c_int),
                ("chroma_scaling_from_luma", C.c_int), ("bit_depth", C.c_int), ("scale", C.c_float),
                ("minValue", C.c_float), ("maxLuma", C.c_float), ("maxChroma", C.c_float),
                ("scaling_shift", C.c_int), ("uv_mult", C.c_float * 3),
                ("uv_mult_luma", C.c_float * 3), ("uv_offset", C.c_float * 3), ("luma_comp", C.c_int)]


class OrcH274Grain(C.Structure):
    _fields_ = [("db", FP), ("db_w", C.c_int), ("seed", C.c_uint32), ("scale", C.c_float),
                ("scale_factor", C.c_float), ("num_entries", C.c_int), ("channel", C.c_int * 3),
                ("num_intervals", C.c_int * 3), ("lower", FP * 3), ("upper", FP * 3),
                ("value", C.POINTER(C.c_uint32) * 3)]


def av1_fixture():
    sh = shim()
    sh.shim_av1_fixture_ptr.restype = C.c_void_p
    return Av1GrainData.from_address(sh.shim_av1_fixture_ptr())


class OrcPeakBuf(C.Structure):
    _fields_ = [("frame_wg_count", C.c_uint32 * 12), ("frame_wg_active", C.c_uint32 * 12),
                ("frame_sum_pq", C.c_uint32 * 12), ("frame_max_pq", C.c_uint32 * 12),
                ("frame_hist", (C.c_uint32 * 64) * 12)]


class OrcPeakParams(C.Structure):
    _fields_ = [("trc", C.c_int), ("csp_min", C.c_float), ("csp_max", C.c_float),
                ("luma_coeffs", C.c_float * 3), ("black_cutoff", C.c_float),
                ("use_histogram", C.c_int), ("wg_w", C.c_int), ("wg_h", C.c_int)]


class OrcPeakState(C.Structure):
    _fields_ = [("avg_pq", C.c_float), ("max_pq", C.c_float)]


class OrcColorMap(C.Structure):
    _fields_ = [("src_trc", C.c_int), ("dst_trc", C.c_int),
                ("src_min", C.c_float), ("src_max", C.c_float),
                ("dst_min", C.c_float), ("dst_max", C.c_float),
                ("src_luma", C.c_float * 3), ("dst_luma", C.c_float * 3),
                ("dst_black_scaled", C.c_int), ("prelinearized", C.c_int),
                ("need_tone_map", C.c_int), ("need_gamut_map", C.c_int),
                ("rgb2lms", C.c_float * 9), ("lms2rgb", C.c_float * 9),
                ("tone_kind", C.c_int), ("tone_lut", FP), ("tone_lut_size", C.c_int),
                ("tone_input_min", C.c_float), ("tone_input_max", C.c_float),
                ("tone_output_min", C.c_float), ("tone_output_max", C.c_float),
                ("tone_exposure", C.c_float),
                ("gamut_lut", FP), ("lut_I", C.c_int), ("lut_C", C.c_int), ("lut_h", C.c_int),
                ("gamut_min_luma", C.c_float), ("gamut_max_luma", C.c_float),
                ("feature_map", C.POINTER(OrcTex)), ("contrast_recovery", C.c_float),
                ("width", C.c_int), ("height", C.c_int)]


class OrcDither(C.Structure):
    _fields_ = [("method", C.c_int), ("lut", FP), ("lut_size", C.c_int), ("new_depth", C.c_int),
                ("temporal", C.c_int), ("frame_index", C.c_uint32), ("gamma", C.c_float)]


def _tex(a, addr=0):
    a = np.ascontiguousarray(a, f32)
    if a.ndim == 2:
        a = a[:, :, None]
    h, w, c = a.shape
    return OrcTex(w, h, c, fp(a), addr), a


class Oracle:
    def __init__(self, lib):
        self.lib = lib

    def num_threads(self):
        return self.lib.orc_num_threads()

    def _ext_out(self, dw, dh, ext):
        """Output buffer for an evaluation extent `ext` = (ew, eh) >= (dw, dh): the overhanging
        invocations of the last compute work groups (see orc_set_eval_extent)."""
        ew, eh = (max(ext[0], dw), max(ext[1], dh)) if ext else (dw, dh)
        self.lib.orc_set_eval_extent(ew if ext else 0, eh if ext else 0)
        return np.zeros((eh, ew, 4), f32)

    def sample_polar(self, src, dw, dh, filt, comp_mask=None, rect=None, scale=1.0,
                     coord_mode=COORD_LITERAL, antiring=None, ext=None):
        """src: HxWxC float32 texel values; filt: dict from T0.filter_generate."""
        t, keep = _tex(src)
        comp_mask = (1 << keep.shape[2]) - 1 if comp_mask is None else comp_mask
        rect = rect or (0, 0, keep.shape[1], keep.shape[0])
        lut = np.ascontiguousarray(filt["weights"], f32)
        p = OrcPolar(fp(lut), lut.size, filt["radius"], filt["radius_zero"],
                     filt["antiring"] if antiring is None else antiring, scale, comp_mask,
                     coord_mode)
        out = self._ext_out(dw, dh, ext)
        self.lib.orc_sample_polar(C.byref(t), OrcRect(*rect), dw, dh, C.byref(p), fp(out))
        self.lib.orc_set_eval_extent(0, 0)
        return out

    def sample_ortho(self, src, dw, dh, filt, pass_, comp_mask=None, rect=None, scale=1.0,
                     coord_mode=COORD_LITERAL, antiring=None, ratio=None, lut=None, address=0):
        t, keep = _tex(src, address)
        comp_mask = (1 << keep.shape[2]) - 1 if comp_mask is None else comp_mask
        rect = rect or (0, 0, keep.shape[1], keep.shape[0])
        use_linear = filt["radius"] == filt["radius_zero"]
        w = np.ascontiguousarray(ortho_lut(filt) if lut is None else lut, f32)
        if ratio is None:
            ratio = dw / (rect[2] - rect[0]) if pass_ == 1 else dh / (rect[3] - rect[1])
        p = OrcOrtho(fp(w), 256, filt["row_size"], filt["row_stride"], int(use_linear),
                     filt["antiring"] if antiring is None else antiring, ratio, scale,
                     comp_mask, pass_, coord_mode)
        out = np.zeros((dh, dw, 4), f32)
        self.lib.orc_sample_ortho(C.byref(t), OrcRect(*rect), dw, dh, C.byref(p), fp(out))
        return out

    def sample_direct(self, src, dw, dh, linear=False, rect=None, scale=1.0,
                      coord_mode=COORD_LITERAL, ext=None, address=0):
        t, keep = _tex(src, address)
        rect = rect or (0, 0, keep.shape[1], keep.shape[0])
        out = self._ext_out(dw, dh, ext)
        self.lib.orc_sample_direct(C.byref(t), OrcRect(*rect), dw, dh, int(linear),
                                   C.c_float(scale), coord_mode, fp(out))
        self.lib.orc_set_eval_extent(0, 0)
        return out

    def sample_fast(self, src, dw, dh, kind, rect=None, scale=1.0, threshold=0.0,
                    coord_mode=COORD_LITERAL, address=0):
        """kind: bicubic / hermite / gaussian / oversample (sampling.c:319-478)."""
        t, keep = _tex(src, address)
        rect = rect or (0, 0, keep.shape[1], keep.shape[0])
        out = self._ext_out(dw, dh, None)
        rx, ry = dw / abs(rect[2] - rect[0]), dh / abs(rect[3] - rect[1])
        self.lib.orc_sample_fast(C.byref(t), OrcRect(*rect), dw, dh,
                                 ["bicubic", "hermite", "gaussian", "oversample"].index(kind),
                                 C.c_float(scale), C.c_float(rx), C.c_float(ry),
                                 C.c_float(min(max(threshold, 0.0), 0.5)), coord_mode, fp(out))
        return out

    def deband(self, src, iterations=1, threshold=3.0, radius=16.0, grain=4.0,
               neutral=(0, 0, 0), scale=1.0, comp_mask=None, frame_index=0, ext=None):
        t, keep = _tex(src)
        comp_mask = (1 << keep.shape[2]) - 1 if comp_mask is None else comp_mask
        p = OrcDeband(iterations, threshold, radius, grain, (C.c_float * 3)(*neutral), scale,
                      comp_mask, frame_index)
        out = self._ext_out(keep.shape[1], keep.shape[0], ext)
        self.lib.orc_deband(C.byref(t), C.byref(p), fp(out))
        self.lib.orc_set_eval_extent(0, 0)
        return out

    def prng(self, x, y, z):
        st = (C.c_uint32 * 3)(x, y, z)
        o = np.zeros(3, f32)
        self.lib.orc_prng(st, fp(o))
        return o, tuple(st)

    def store_round(self, a, kind, bits=16):
        a = np.ascontiguousarray(a, f32).copy()
        self.lib.orc_store_round(fp(a), C.c_size_t(a.size), kind, bits)
        return a

    def store_unorm(self, a, bits):
        a = np.ascontiguousarray(a, f32)
        o = np.zeros(a.shape, np.uint16)
        self.lib.orc_store_unorm(fp(a), C.c_size_t(a.size), bits, o.ctypes.data_as(C.c_void_p))
        return o

    def store_half(self, a):
        a = np.ascontiguousarray(a, f32)
        o = np.zeros(a.shape, np.uint16)
        self.lib.orc_store_half(fp(a), C.c_size_t(a.size), o.ctypes.data_as(C.c_void_p))
        return o.view(np.float16)

    def apply_matrix(self, rgba, m, c=None):
        rgba = np.ascontiguousarray(rgba, f32).copy()
        m = np.ascontiguousarray(m, f32)
        cc = None if c is None else fp(np.ascontiguousarray(c, f32))
        self.lib.orc_apply_matrix(fp(rgba), C.c_size_t(rgba.size // 4), fp(m), cc)
        return rgba

    def scale(self, rgba, s, with_alpha=True):
        rgba = np.ascontiguousarray(rgba, f32).copy()
        self.lib.orc_scale(fp(rgba), C.c_size_t(rgba.size // 4), C.c_float(s), int(with_alpha))
        return rgba

    def _trc(self, fn, rgba, trc, csp_min, csp_max, luma, *extra):
        rgba = np.ascontiguousarray(rgba, f32).copy()
        luma = np.ascontiguousarray(luma if luma is not None else [0.2126, 0.7152, 0.0722], f32)
        fn(fp(rgba), C.c_size_t(rgba.size // 4), trc, C.c_float(csp_min), C.c_float(csp_max),
           fp(luma), *extra)
        return rgba

    def linearize(self, rgba, trc, csp_min=0.0, csp_max=1.0, luma=None):
        return self._trc(self.lib.orc_linearize, rgba, trc, csp_min, csp_max, luma)

    def delinearize(self, rgba, trc, csp_min=0.0, csp_max=1.0, luma=None, black_scaled=True):
        return self._trc(self.lib.orc_delinearize, rgba, trc, csp_min, csp_max, luma,
                         int(black_scaled))

    def decode_special(self, rgba, sys):
        rgba = np.ascontiguousarray(rgba, f32).copy()
        self.lib.orc_decode_special(fp(rgba), C.c_size_t(rgba.size // 4), sys)
        return rgba

    def encode_special(self, rgba, sys):
        rgba = np.ascontiguousarray(rgba, f32).copy()
        self.lib.orc_encode_special(fp(rgba), C.c_size_t(rgba.size // 4), sys)
        return rgba

    def av1_grain(self, plane, luma, params):
        """plane: HxWxC float32 texel values (unorm/65535), luma likewise or None."""
        t, keep = _tex(plane)
        tl, keep2 = _tex(luma) if luma is not None else (None, None)
        out = np.zeros((keep.shape[0], keep.shape[1], 4), f32)
        self.lib.orc_av1_grain(C.byref(t), C.byref(tl) if tl is not None else None, C.byref(params), fp(out))
        return out

    def h274_grain(self, plane, params):
        """plane: HxWxC float32 texel values; returns HxWx4."""
        t, keep = _tex(plane)
        out = np.zeros((keep.shape[0], keep.shape[1], 4), f32)
        self.lib.orc_h274_grain(C.byref(t), C.byref(params), fp(out))
        return out

    def extract_features(self, rgba, m):
        """rgba: linear light; m: rgb2lms (9 floats). Returns (I, 0, 0, 1)."""
        rgba = np.ascontiguousarray(rgba, f32).copy()
        m = np.ascontiguousarray(m, f32)
        self.lib.orc_extract_features(fp(rgba), C.c_size_t(rgba.size // 4), fp(m))
        return rgba

    def deinterlace(self, cur, prev=None, nxt=None, field=1, first_field=0, algo=3, skip_spatial_check=False,
                    comp_mask=0):
        """pl_shader_deinterlace on float textures (h, w, comps); returns (h, w, 4)."""
        tc, kc = _tex(cur)
        tp, kp = _tex(prev) if prev is not None else (None, None)
        tn, kn = _tex(nxt) if nxt is not None else (None, None)
        out = np.zeros(kc.shape[:2] + (4,), f32)
        self.lib.orc_deinterlace(C.byref(tc), C.byref(tp) if tp is not None else None,
                                 C.byref(tn) if tn is not None else None, field, first_field, algo,
                                 int(skip_spatial_check), comp_mask, fp(out))
        return out

    def mix_accumulate(self, acc, rgba, w):
        acc = np.ascontiguousarray(acc, f32).copy()
        rgba = np.ascontiguousarray(rgba, f32)
        self.lib.orc_mix_accumulate(fp(acc), fp(rgba), C.c_size_t(rgba.size // 4), C.c_float(w))
        return acc

    def custom_lut(self, rgba, data, size, shaper_in=None, shaper_out=None):
        """pl_shader_custom_lut: data = RGB triples (red fastest); size = (n, 0, 0) or (r, g, b)."""
        rgba = np.ascontiguousarray(rgba, f32).copy()
        d = np.ascontiguousarray(data, f32)
        sz = (C.c_int * 3)(*size)
        si = None if shaper_in is None else fp(np.ascontiguousarray(shaper_in, f32))
        so = None if shaper_out is None else fp(np.ascontiguousarray(shaper_out, f32))
        self.lib.orc_custom_lut(fp(rgba), C.c_size_t(rgba.size // 4), fp(d), sz, si, so)
        return rgba

    def dovi_reshape(self, rgba, meta):
        rgba = np.ascontiguousarray(rgba, f32).copy()
        comps = (OrcDoviComp * 3)()
        for c in range(3):
            s, d = meta.comp[c], comps[c]
            d.num_pivots = s.num_pivots
            for i in range(9):
                d.pivots[i] = s.pivots[i]
            for i in range(8):
                d.method[i], d.mmr_order[i], d.mmr_constant[i] = s.method[i], s.mmr_order[i], s.mmr_constant[i]
                for k in range(3):
                    d.poly_coeffs[i][k] = s.poly_coeffs[i][k]
                    for j in range(7):
                        d.mmr_coeffs[i][k][j] = s.mmr_coeffs[i][k][j]
        self.lib.orc_dovi_reshape(fp(rgba), C.c_size_t(rgba.size // 4), comps)
        return rgba

    def dovi_lms(self, rgba, lms2rgb):
        rgba = np.ascontiguousarray(rgba, f32).copy()
        m = np.ascontiguousarray(lms2rgb, f32)
        self.lib.orc_dovi_lms(fp(rgba), C.c_size_t(rgba.size // 4), fp(m))
        return rgba

    def sigmoidize(self, rgba, center=0.75, slope=6.5, inverse=False):
        rgba = np.ascontiguousarray(rgba, f32).copy()
        fn = self.lib.orc_unsigmoidize if inverse else self.lib.orc_sigmoidize
        fn(fp(rgba), C.c_size_t(rgba.size // 4), C.c_float(center), C.c_float(slope))
        return rgba

    def detect_peak(self, rgba, trc, csp_min, csp_max, luma, black_cutoff=1.0,
                    use_histogram=False, wg=(16, 16)):
        rgba = np.ascontiguousarray(rgba, f32)
        h, w = rgba.shape[:2]
        p = OrcPeakParams(trc, csp_min, csp_max, (C.c_float * 3)(*luma), black_cutoff,
                          int(use_histogram), wg[0], wg[1])
        buf = OrcPeakBuf()
        self.lib.orc_detect_peak(fp(rgba), w, h, C.byref(p), C.byref(buf))
        return buf

    def peak_update(self, state, buf, smoothing=20.0, low=1.0, high=3.0, percentile=100.0):
        st = OrcPeakState(*state)
        used = self.lib.orc_peak_update(C.byref(st), C.byref(buf), C.c_float(smoothing),
                                        C.c_float(low), C.c_float(high), C.c_float(percentile))
        return (st.avg_pq, st.max_pq), bool(used)

    def color_map(self, rgba, params):
        rgba = np.ascontiguousarray(rgba, f32).copy()
        self.lib.orc_color_map(fp(rgba), C.c_size_t(rgba.size // 4), C.byref(params))
        return rgba

    def error_diffusion(self, rgba, new_depth, kernel):
        """kernel: dict from T0.error_diffusion_kernel"""
        src = np.ascontiguousarray(rgba, f32)
        h, w = src.shape[:2]
        out = np.zeros_like(src)
        pat = np.ascontiguousarray(kernel["pattern"], np.int32)
        self.lib.orc_error_diffusion(fp(src), fp(out), w, h, new_depth, kernel["shift"],
                                     kernel["divisor"], ip(pat))
        return out

    def dither(self, rgba, new_depth, method=0, lut=None, temporal=False, frame_index=0,
               gamma=1.0):
        rgba = np.ascontiguousarray(rgba, f32).copy()
        h, w = rgba.shape[:2]
        l = np.ascontiguousarray(lut, f32) if lut is not None else None
        p = OrcDither(method, fp(l) if l is not None else None,
                      l.shape[0] if l is not None else 0, new_depth, int(temporal), frame_index,
                      gamma)
        self.lib.orc_dither(fp(rgba), w, h, C.byref(p))
        return rgba


def ortho_lut(filt):
    """fill_ortho_lut (src/shaders/sampling.c:914-948) applied to pl_filter weights."""
    w = filt["weights"].reshape(256, filt["row_stride"]).copy()
    if filt["radius"] != filt["radius_zero"]:
        return w
    out = np.zeros_like(w)
    for k in range(0, filt["row_size"], 2):
        w0, w1 = w[:, k], w[:, k + 1]
        out[:, k] = w0 + w1
        out[:, k + 1] = w1 / (w0 + w1)
    for k in range(filt["row_size"], filt["row_stride"]):
        out[:, k] = out[:, k - 4] if k >= 4 else 0
    return out


# ----------------------------------------------------------------------------------------
# synthetic frames (the bench.c ring pattern + seeded noise)
# ----------------------------------------------------------------------------------------
def rings(w, h, comps=3):
    """0.5*sin(f*r^2)+0.5 concentric rings, per-component frequency (src/tests/bench.c:32-52)."""
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    cx, cy = w / 2.0, h / 2.0
    r2 = (x - cx) ** 2 + (y - cy) ** 2
    phi = (1 + 5 ** 0.5) / 2
    f = 0.1 * np.pi / np.hypot(cx, cy)
    out = []
    for _ in range(comps):
        out.append(0.5 * np.sin(f * r2) + 0.5)
        f /= phi
    return np.stack(out, -1)


def yuv420p10(w, h, seed=0, kind="rings"):
    """Y (h,w), U, V (h/2,w/2) uint16 planes holding 10-bit limited-range code values."""
    rng = np.random.default_rng(seed)
    if kind == "noise":
        y = rng.integers(64, 941, (h, w))
        u = rng.integers(64, 961, (h // 2, w // 2))
        v = rng.integers(64, 961, (h // 2, w // 2))
    else:
        r = rings(w, h, 3)
        y = 64 + r[..., 0] * (940 - 64)
        c = rings(w // 2, h // 2, 3)
        u = 64 + c[..., 1] * (960 - 64)
        v = 64 + c[..., 2] * (960 - 64)
        # a little noise so that debanding has something to decide on
        y = y + rng.integers(-2, 3, y.shape)
        u = u + rng.integers(-2, 3, u.shape)
        v = v + rng.integers(-2, 3, v.shape)
    clip = lambda a: np.clip(np.rint(a), 0, 1023).astype(np.uint16)
    return clip(y), clip(u), clip(v)


def unorm16_to_float(a):
    return (a.astype(f32) / f32(65535.0)).astype(f32)


def ulp_diff_half(a, b):
    """Distance in fp16 ULPs between two float16 arrays (monotone integer mapping)."""
    def key(x):
        i = x.view(np.int16).astype(np.int32)
        return np.where(i < 0, -(i & 0x7FFF), i)
    return np.abs(key(a) - key(b))
