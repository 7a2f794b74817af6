"""CPU restatement of the reference's per-frame pass graph (src/renderer.c pl_render_image)
for planar YCbCr 4:2:0 sources and one packed RGB target, composed from the oracle's per-stage
functions (oracle/pl_oracle*.c). It is written independently of libplacebo_b200/csrc/renderer.c
(numpy, different structure) so that the two planners check each other, and it is also the
"port" CPU baseline that bench.py times.

Pass graph (file:line in /root/reference/src/renderer.c):
  merge U+V into one rg16 texture (exact copy)                               :1617-1699
  per plane: deband (:1318) -> chroma stored to an rg16hf FBO (:1861) -> plane scaler (:1864)
  merged pass: planes -> colour, decode_color (:1942)                        :1788-1957
  main scaler: [peak detect] [linearize] [sigmoid] -> rgba16hf FBO -> polar  :1964-2087
  colour map: [rgba16hf FBO when same-frame peak] -> color_map_ex            :2157-2279
  output: encode_color, dither, 1/scale, store                               :2586-2964
"""
import ctypes as C
import os

import numpy as np

import plbtest as T

f32 = np.float32


def _filter_for(t0, name, rx, ry, no_widening=False):
    """pl_shader_sample_polar's filter config: blur widened when downscaling (sampling.c:608-616)"""
    inv = f32(1.0) / f32(min(rx, ry))
    inv = max(inv, f32(1.0))
    if no_widening:
        inv = f32(1.0)
    base = t0.filter_generate(name)
    blur = f32(base["blur"] if base["blur"] else 1.0) * f32(inv)
    return t0.filter_generate(name, blur=float(blur)) if blur != 1.0 else base


def luma_coeffs(t0, prim):
    return t0.matrix(0, t0.raw_primaries(prim))[3:6].copy()


class Frame:
    def __init__(self, planes, sys, levels=T.LEVELS["LIMITED"], sample_depth=16, color_depth=10,
                 bit_shift=0, csp=None, chroma_loc="left", dovi=None):
        self.planes = planes
        self.dovi = dovi  # Dolby Vision: variant of the reference's dovi_meta fixture (0 / 1)
        self.sys, self.levels = sys, levels
        self.bits = (sample_depth, color_depth, bit_shift)
        self.csp = csp if csp is not None else T.csp()
        self.chroma_loc = chroma_loc


class Target:
    """`planes`: None for one packed plane of format `fmt`, else a list of (format, component
    indices, (sub_x, sub_y)) — e.g. yuv420p16 = [("r16", (0,), (1, 1)), ("r16", (1,), (2, 2)),
    ("r16", (2,), (2, 2))], P016 = [("r16", (0,), (1, 1)), ("rg16", (1, 2), (2, 2))]; subsampled
    planes are sited at `chroma_loc`."""
    def __init__(self, w, h, fmt="rgba16hf", sys=T.SYS["RGB"], levels=T.LEVELS["FULL"],
                 sample_depth=0, color_depth=0, csp=None, planes=None, chroma_loc="left"):
        self.w, self.h, self.fmt = w, h, fmt
        self.sys, self.levels = sys, levels
        self.bits = (sample_depth, color_depth, 0)
        self.csp = csp if csp is not None else T.csp()
        self.planes, self.chroma_loc = planes, chroma_loc

    def plane_size(self, i):
        sx, sy = self.planes[i][2]
        return (self.w + sx - 1) // sx, (self.h + sy - 1) // sy


CHROMA_SHIFT = {"left": (-0.5, 0.0), "center": (0.0, 0.0), "top_left": (-0.5, -0.5)}
CHROMA_ENUM = {"left": 1, "center": 2, "top_left": 3}


def fix_bits(fmt_depth, bits):
    """fix_frame (renderer.c:3131-3160) for UNORM textures"""
    sd, cd, sh = bits
    if not sd and fmt_depth:
        sd = fmt_depth
        cd = cd or sd
        cd = min(cd, sd)
        sh += sd - cd
    return sd, cd, sh


# ---- AV1 film grain: LUTs come from the real reference shader (oracle/harvest.c) ---------------
HARVEST_SO = os.path.join(T.ROOT, "oracle", "_ref", "libplref_harvest.so")


def av1_harvest(plane, w, h, seed, depth, levels=1, sys_=2):
    """LUT descriptors uploaded by pl_shader_film_grain for the reference's fixture.
    plane 0: luma r16 w x h; plane 1: merged chroma rg16 (w/2 x h/2, luma tex w x h)."""
    import tempfile
    hv = C.CDLL(HARVEST_SO)
    d = tempfile.mkdtemp()
    assert hv.hvs_av1_grain(d.encode(), w, h, plane, seed, 1, depth, levels, sys_) == 0
    descs = {}
    for line in open(os.path.join(d, "meta.txt")):
        f = line.split()
        if f and f[0] == "desc" and f[-1] == "1":
            descs[int(f[1])] = (f[4], int(f[5]), int(f[6]), os.path.join(d, f"desc_{f[1]}.bin"))
    return descs


def av1_oracle_params(descs, chroma, depth, overlap, keep, tex_scale):
    """struct orc_av1_grain_params for the fixture; host constants as film_grain_av1.c:860-880
    (limited range; `bits` there is the normalised repr's depth = 16)."""
    fx = T.av1_fixture()
    p = T.OrcAv1Grain()
    if chroma:
        uv = np.fromfile(descs[0][3], f32).reshape(-1, 2)
        g = [np.ascontiguousarray(uv[:, 0]), np.ascontiguousarray(uv[:, 1])]
        sc = [np.fromfile(descs[2][3], f32), np.fromfile(descs[3][3], f32)]
        p.grain_w = p.grain_h = 32
        p.sub_x = p.sub_y = 1
    else:
        g = [np.fromfile(descs[0][3], f32)]
        sc = [np.fromfile(descs[2][3], f32)]
        p.grain_w = p.grain_h = 64
    off = np.fromfile(descs[1][3], np.uint32)
    keep += g + sc + [off]
    for i in range(len(g)):
        p.grain[i] = T.fp(g[i])
        p.scaling[i] = T.fp(sc[i])
        p.is_luma[i] = 0 if chroma else 1
        p.uv_mult[i] = fx.uv_mult[i] / 64.0
        p.uv_mult_luma[i] = fx.uv_mult_luma[i] / 64.0
        p.uv_offset[i] = float(int(fx.uv_offset[i]) << (16 - 8)) / ((1 << depth) - 1)
    p.offsets = off.ctypes.data_as(C.POINTER(C.c_uint32))
    p.offsets_w, p.offsets_h = descs[1][1], descs[1][2]
    p.num_comps = len(g)
    p.overlap = overlap
    p.chroma_scaling_from_luma = int(fx.chroma_scaling_from_luma)
    p.bit_depth = depth
    p.scale = float(tex_scale)
    out_scale = float(f32(65536 / 65535.0))
    p.minValue = float(f32(16 / 256.0 * out_scale))
    p.maxLuma = float(f32(235 / 256.0 * out_scale))
    p.maxChroma = float(f32(240 / 256.0 * out_scale))
    p.scaling_shift = fx.scaling_shift
    p.luma_comp = 0
    return p


def render(img, tgt, *, upscaler=None, plane_upscaler=None, deband=False, peak=0,
           allow_delayed=False, dither=0, cmap_hq=False, frame_index=0, peak_state=None,
           coord_mode=T.COORD_EXACT, tone_fn=None, gamut_fn=None, stats=None,
           error_diffusion=None, film_grain=None, contrast_recovery=0.0, contrast_smoothness=3.5,
           luts=None, deinterlace=None, _return_color=False, _start_color=None):
    """Returns (output array in target storage type, new peak state). `peak_state` carries the
    smoothed (avg, max) between frames like sh_color_map_obj.peak does.
    luts: {"params": (type, data, size), "image": (...), "target": (...)} with type in
    "native" / "normalized" / "conversion" (custom LUTs; the combinations the renderer tests use:
    params native / conversion, image normalized, target native)."""
    luts = luts or {}

    def apply_lut(which, c):
        _, data, size = luts[which]
        return orc.custom_lut(c, data, size)

    plut = luts.get("params")
    if plut and plut[0] == "conversion":
        peak = 0                      # LUT handles tone mapping (renderer.c:1224, :2105)
        contrast_recovery = 0.0
    t0, orc = T.t0(), T.oracle()
    Y, U, V = img.planes
    if deinterlace is not None:
        # (field, first_field, algo, prev Frame or None, next Frame or None): every plane is
        # deinterlaced into a texture of its own format (r16) first, renderer.c:1591-1612
        field, first, algo, pimg, nimg = deinterlace
        fix = []
        for i, pl in enumerate((Y, U, V)):
            f = lambda a: T.unorm16_to_float(a)[:, :, None]
            o = orc.deinterlace(f(pl), f(pimg.planes[i]) if pimg is not None else None,
                                f(nimg.planes[i]) if nimg is not None else None,
                                field=field, first_field=first, algo=algo, comp_mask=1)
            fix.append(orc.store_unorm(np.ascontiguousarray(o[..., :1]), 16)[..., 0])
        Y, U, V = fix
    H, W = Y.shape
    ch, cw = U.shape
    plane_up = plane_upscaler or upscaler

    # ---- colour spaces / representation fix-ups (pass_fix_frames :3195-3233)
    src_csp = img.csp.copy()
    if not src_csp[0]:
        src_csp[0] = T.PRIM["BT_709"] if (W >= 1280 or H > 576) else (
            T.PRIM["BT_601_625"] if H == 576 else T.PRIM["BT_601_525"] if H in (480, 486) else T.PRIM["BT_709"])
    src_csp, dst_csp = t0.csp_infer_map(src_csp, tgt.csp)
    sbits = fix_bits(16, img.bits)
    tdepth = {"rgba8": 8, "rgba16": 16}.get(tgt.fmt, 0)
    tbits = fix_bits(tdepth, tgt.bits)

    # pl_color_repr_normalize depends on the bit depths only (a Dolby Vision repr needs its
    # metadata for anything else, so ask with a neutral system)
    norm_sys = T.SYS["BT_2020_NC"] if img.dovi is not None else img.sys
    norm = t0.repr_decode(norm_sys, img.levels, *sbits)["scale"]  # pl_color_repr_normalize
    ycbcr = img.sys not in (T.SYS["UNKNOWN"], T.SYS["RGB"], T.SYS["XYZ"])
    limited = (img.levels or (T.LEVELS["LIMITED"] if ycbcr else T.LEVELS["FULL"])) == T.LEVELS["LIMITED"]
    out_scale = f32(2.0 ** sbits[0]) / (f32(2.0 ** sbits[0]) - f32(1.0))
    neutral_luma = f32(16 / 256.0) * out_scale if limited else f32(0.0)
    neutral_chroma = f32(0.5) * out_scale if ycbcr else neutral_luma

    Yf = T.unorm16_to_float(Y)[:, :, None]
    UVf = np.stack([T.unorm16_to_float(U), T.unorm16_to_float(V)], -1)  # merged rg16: exact

    # ---- B. per plane
    grain = f32(4.0) / (f32(src_csp[11]) / f32(203.0))  # dparams.grain /= max_luma / SDR_WHITE
    sx, sy = CHROMA_SHIFT[img.chroma_loc]
    rrx, rry = cw / W, ch / H
    crect = ((0 - sx) * rrx, (0 - sy) * rry, (W - sx) * rrx, (H - sy) * rry)
    # The merged pass runs as whole work groups: rigid 32x32 if it contains a polar compute
    # sampler (sh_subpass -> sh_try_compute, shaders.c:677-686), else the flexible 16x16 of
    # detect_peak. Invocations past the image edge still execute (only the store is masked) and
    # peak detection counts them, so everything merged is evaluated on the padded extent.
    merged_has_polar = bool(plane_up) and (cw != W or ch != H) and plane_up.startswith("ewa")
    h274 = film_grain is not None and film_grain[1] == "h274"
    # A merged grain sub-shader brings its own rigid geometry: 32x32 for AV1 luma
    # (film_grain_av1.c:693), 8x8 for H.274 (film_grain_h274.c:164) -- and a rigid 8x8 pass
    # cannot host the polar sampler's 32x32 (sh_subpass fails, renderer.c:1866-1875): the
    # scaled chroma goes through an FBO instead.
    wg = (8, 8) if h274 else (32, 32) if (merged_has_polar or film_grain is not None) else (16, 16)
    ext = (-(-W // wg[0]) * wg[0], -(-H // wg[1]) * wg[1])
    if deband:
        luma = orc.deband(Yf, 1, 3.0, 16.0, float(grain), (neutral_luma, 0, 0), float(norm), 1,
                          frame_index, ext=ext)
        luma_v = luma[..., 0]
        cdb = orc.deband(UVf, 1, 3.0, 16.0, float(grain), (neutral_chroma, neutral_chroma, 0),
                         float(norm), 3, frame_index)
        ctex = orc.store_round(cdb[..., :2], 0)  # rg16hf FBO
        cscale = 1.0
    else:
        luma_v = orc.sample_direct(Yf, W, H, scale=float(norm), coord_mode=coord_mode, ext=ext)[..., 0]
        ctex, cscale = UVf, float(norm)

    # ---- film grain (AV1), per plane after the optional debanding (renderer.c:1355-1429)
    if film_grain is not None and film_grain[1] == "h274":
        # the reference's H.274 fixture models luma only: the chroma planes need no pass
        # (pl_needs_film_grain, film_grain_h274.c:127-143)
        from test_grain_tables import harvest_h274
        db, _ = harvest_h274()
        if deband:
            ltex, lscale = orc.store_round(luma[:H, :W, :1], 0), 1.0
        else:
            ltex, lscale = Yf, float(norm)
        hp = T.OrcH274Grain()
        lo, hi = np.array([10 / 255.0], f32), np.array([250 / 255.0], f32)
        val = np.array([16 << 16 | 10 << 8 | 12], np.uint32)
        hp.db, hp.db_w, hp.seed = T.fp(db), 832, film_grain[0]
        hp.scale, hp.scale_factor = lscale, 2.0 ** -8
        hp.num_entries, hp.channel[0], hp.num_intervals[0] = 1, 0, 1
        hp.lower[0], hp.upper[0] = T.fp(lo), T.fp(hi)
        hp.value[0] = val.ctypes.data_as(C.POINTER(C.c_uint32))
        lg = orc.h274_grain(ltex, hp)[..., 0]
        luma_v = np.pad(lg, ((0, ext[1] - H), (0, ext[0] - W)), mode="edge")
    elif film_grain is not None:
        seed, overlap = film_grain
        depth = img.bits[1]
        keep = []
        if deband:
            ltex, lscale = orc.store_round(luma[:H, :W, :1], 0), 1.0   # r16hf FBO of the debanded plane
        else:
            ltex, lscale = Yf, float(norm)
        pl = av1_oracle_params(av1_harvest(0, W, H, seed, depth, img.levels, img.sys), 0, depth,
                               overlap, keep, lscale)
        lg = orc.av1_grain(ltex, None, pl)[..., 0]
        luma_v = np.pad(lg, ((0, ext[1] - H), (0, ext[0] - W)), mode="edge")
        pc = av1_oracle_params(av1_harvest(1, W, H, seed, depth, img.levels, img.sys), 1, depth,
                               overlap, keep, cscale)
        # quirk kept from the reference: the luma fetched for chroma scaling is the ORIGINAL
        # (un-normalised) plane times this plane's texture scale
        cg = orc.av1_grain(ctex, Yf, pc)
        ctex, cscale = orc.store_round(cg[..., :2], 0), 1.0                 # rg16hf FBO

    # plane scaler (SAMPLER_PLANE)
    crect_f = tuple(f32(v) for v in crect)
    if plane_up and (cw != W or ch != H) and h274 and merged_has_polar:
        rx, ry = W / (crect[2] - crect[0]), H / (crect[3] - crect[1])
        filt = _filter_for(t0, plane_up, rx, ry)
        cfbo = orc.sample_polar(ctex, W, H, filt, comp_mask=3, rect=crect_f, scale=cscale,
                                coord_mode=coord_mode)
        cfbo = orc.store_round(cfbo[..., :2], 0)                            # rg16hf FBO
        chroma = np.pad(cfbo, ((0, ext[1] - H), (0, ext[0] - W), (0, 0)), mode="edge")
    elif plane_up and (cw != W or ch != H):
        rx, ry = W / (crect[2] - crect[0]), H / (crect[3] - crect[1])
        filt = _filter_for(t0, plane_up, rx, ry)
        chroma = orc.sample_polar(ctex, W, H, filt, comp_mask=3, rect=crect_f, scale=cscale,
                                  coord_mode=coord_mode, ext=ext)
    else:
        chroma = orc.sample_direct(ctex, W, H, linear=(cw != W or ch != H), rect=crect_f,
                                   scale=cscale, coord_mode=coord_mode, ext=ext)

    # ---- C. merge + decode (on the padded extent; cropped after the peak measurement)
    color = np.empty((ext[1], ext[0], 4), f32)
    color[..., 0] = luma_v
    color[..., 1] = chroma[..., 0]
    color[..., 2] = chroma[..., 1]
    color[..., 3] = 1.0
    if img.dovi is not None:
        # pl_shader_decode_color for PL_COLOR_SYSTEM_DOLBYVISION (colorspace.c:285-310, 392-422):
        # residual scale of the normalised repr, reshape, RPU matrix + offsets, PQ EOTF, LMS -> RGB,
        # PQ OETF
        meta = T.dovi_fixture(img.dovi)
        d = t0.repr_decode_dovi(sbits[0], sbits[0], list(meta.nonlinear_offset), list(meta.nonlinear),
                                list(meta.linear))
        if float(d["scale"]) != 1.0:
            color = orc.scale(color, float(d["scale"]), with_alpha=False)
        color = orc.dovi_reshape(color, meta)
        color = orc.apply_matrix(color, d["m"], d["c"])
        color = orc.dovi_lms(color, d["lms2rgb"])
    else:
        dec = t0.repr_decode(img.sys, img.levels, sbits[0], sbits[0], 0)  # repr already normalised
        color = orc.apply_matrix(color, dec["m"], dec["c"])
        if img.sys in (T.SYS["BT_2020_C"], T.SYS["BT_2100_PQ"], T.SYS["BT_2100_HLG"]):
            color = orc.decode_special(color, img.sys)

    if "image" in luts:
        assert luts["image"][0] == "normalized"   # RGB -> RGB, after decoding (renderer.c:1945)
        color = apply_lut("image", color)

    src_trc = int(src_csp[1])
    src_hdr = src_csp[11] > 203.0 or src_trc in (T.TRC["PQ"], T.TRC["HLG"], T.TRC["V_LOG"],
                                                 T.TRC["S_LOG1"], T.TRC["S_LOG2"], T.TRC["SCRGB"])
    src_luma = luma_coeffs(t0, int(src_csp[0]))
    smin, smax, _ = t0.nominal_luma(src_csp, T.META["HDR10"], T.HDR["NORM"])

    # ---- peak detection decision (hdr_update_peak :1183-1250)
    def want_peak():
        if not peak or not src_hdr:
            return False
        max_peak = {T.TRC["PQ"]: 10000.0, T.TRC["SCRGB"]: 10000.0}.get(src_trc, 203.0)
        if src_trc == T.TRC["HLG"]:
            max_peak = src_csp[11]
        return not (max_peak <= dst_csp[11] + 1e-6) and not src_csp[19]

    new_state = peak_state
    def run_peak(image, wg):
        nonlocal new_state
        hq = peak == 2
        buf = orc.detect_peak(image, src_trc, smin, smax, src_luma, 1.0, hq, wg)
        st, used = orc.peak_update(peak_state or (0.0, 0.0), buf, 20.0, 1.0, 3.0,
                                   99.995 if hq else 100.0)
        new_state = st
        if stats is not None:
            stats["peak_buf"] = buf

    # ---- D. main scaler
    out_w, out_h = tgt.w, tgt.h
    up = out_w > W or out_h > H
    down = out_w < W or out_h < H
    scaling = up or down
    measured = False
    if scaling and up and want_peak():
        run_peak(color, wg)
        measured = True
    if scaling:
        main = upscaler if up else None
        img_fbo = orc.store_round(color[:H, :W], 0)  # rgba16hf FBO at source size
        if main and main.startswith("ewa"):
            filt = _filter_for(t0, main, out_w / W, out_h / H)
            wg = (32, 32)
            ext = (-(-out_w // 32) * 32, -(-out_h // 32) * 32)
            color = orc.sample_polar(img_fbo, out_w, out_h, filt, comp_mask=7, coord_mode=coord_mode,
                                     ext=ext)
        else:
            wg = (16, 16)
            ext = (-(-out_w // 16) * 16, -(-out_h // 16) * 16)
            color = orc.sample_direct(img_fbo, out_w, out_h, linear=True, coord_mode=coord_mode,
                                      ext=ext)
    in_fbo = False
    lut_done = False
    if not measured and want_peak():
        run_peak(color, wg)
        measured = True
        if not allow_delayed:
            if plut:  # the LUT is recorded into the image shader before it becomes the FBO
                assert plut[0] == "native"
                color = apply_lut("params", color)
                lut_done = True
            color = orc.store_round(color, 0)  # need_peak_fbo: rgba16hf at output size
            in_fbo = True
    color = np.ascontiguousarray(color[:out_h, :out_w])
    if plut and not lut_done:
        # native: image colour space in and out (both colour maps around it are no-ops);
        # conversion: the LUT output is the target colour
        color = apply_lut("params", color)

    # ---- contrast recovery feature map (get_feature_map, renderer.c:2089-2154)
    feature = None
    if (contrast_recovery and contrast_smoothness > 1 and src_hdr
            and not (src_csp[11] <= dst_csp[11] + 1e-6)):
        if not in_fbo:
            color = orc.store_round(color, 0)  # img_tex: the image goes through its rgba16hf FBO
        lin = orc.linearize(color, src_trc, smin, smax, src_luma)
        feat = orc.extract_features(lin, t0.matrix(2, t0.raw_primaries(int(src_csp[0]))))
        inter = orc.store_round(np.ascontiguousarray(feat[..., :1]), 0)  # r16hf FBO, full size
        cr_w = int(np.ceil(out_w / f32(contrast_smoothness)))
        cr_h = int(np.ceil(out_h / f32(contrast_smoothness)))
        # bicubic down-scaling, kernel widened by the scaling ratio, mirrored borders: vertical
        # pass into an FBO, then the horizontal pass (dispatch_sampler, renderer.c:742-768)
        fv = t0.filter_generate("bicubic", blur=out_h / cr_h, polar=False)
        v = orc.sample_ortho(inter, out_w, cr_h, fv, 0, comp_mask=1, coord_mode=coord_mode,
                             ratio=cr_h / out_h, address=2)
        v = orc.store_round(np.ascontiguousarray(v[..., :1]), 0)
        fh = t0.filter_generate("bicubic", blur=out_w / cr_w, polar=False)
        hh = orc.sample_ortho(v, cr_w, cr_h, fh, 1, comp_mask=1, coord_mode=coord_mode,
                              ratio=cr_w / out_w, address=2)
        feature = orc.store_round(np.ascontiguousarray(hh[..., :1]), 0)

    # ---- colour mapping (pl_shader_color_map_ex host logic, colorspace.c:1618-1790)
    src2 = src_csp.copy()
    use = new_state if (measured and not allow_delayed) else peak_state
    if peak and use and use[0]:
        src2[19], src2[18] = use[0], use[1]
    cm = color_map_params(t0, src2, dst_csp, hq=cmap_hq, tone_fn=tone_fn, gamut_fn=gamut_fn)
    if plut and plut[0] == "conversion":
        cm = None
    if cm is not None:
        p = cm["params"]
        if (feature is not None and p.need_tone_map
                and float(p.tone_input_max) >= float(p.tone_output_max)):
            ftex, fkeep = T._tex(feature)
            p.feature_map = C.pointer(ftex)
            p.contrast_recovery = float(contrast_recovery)
            p.width, p.height = out_w, out_h
        color = orc.color_map(color, p)

    if _return_color:      # the image in the target colour space, before the output stage
        return color, new_state
    if _start_color is not None:
        color = _start_color
    # ---- output: encode, dither, scale, store
    enc_skip = (tgt.sys in (T.SYS["UNKNOWN"], T.SYS["RGB"]) and tgt.levels in (0, T.LEVELS["FULL"])
                and (not tbits[0] or not tbits[1] or tbits[0] == tbits[1]) and not tbits[2])
    tnorm = t0.repr_decode(tgt.sys, tgt.levels, *tbits)["scale"]
    if not enc_skip:
        enc = t0.repr_encode(tgt.sys, tgt.levels, tbits[0], tbits[0], 0)
        color = orc.apply_matrix(color, enc["m"], enc["c"])
    if "target" in luts:
        assert luts["target"][0] == "native"      # after encoding (renderer.c:2784-2785)
        color = apply_lut("target", color)
    depth = tbits[1]
    if tgt.planes:
        return planar_output(orc, t0, color, tgt, tnorm, depth, dither, int(dst_csp[1]), frame_index,
                             coord_mode), new_state
    if depth and depth < 16 and error_diffusion:
        # pass_error_diffusion (renderer.c:2282-2343): image -> rgba16hf FBO -> single work
        # group diffusion -> rgba16hf FBO -> sampled 1:1
        k = t0.error_diffusion_kernel(error_diffusion)
        fbo = orc.store_round(color, 0)  # float values after rounding to half
        color = orc.store_round(orc.error_diffusion(fbo, depth, k), 0)
    elif depth and depth < 16 and dither:
        lut = t0.dither_matrix(0 if dither == 1 else 1, 64) if dither in (1, 2) else None
        method = {1: 0, 2: 1, 3: 2, 4: 3}[dither]
        gamma = approx_gamma(int(dst_csp[1]))
        color = orc.dither(color, depth, method, lut, False, frame_index, gamma)
    color = orc.scale(color, float(f32(1.0 / tnorm)), True)
    color[..., 3] = 1.0  # swizzle_color leaves alpha = 1 for opaque images
    return store(orc, color, tgt.fmt), new_state


def mix_weights(timestamps, vsync, mixer, t0):
    """Weights of pl_render_image_mix (renderer.c:3661-3757). Returns [(frame index, weight)]."""
    n = len(timestamps)
    ref = 0
    for i in range(1, n):  # pl_frame_mix_nearest
        if abs(timestamps[i]) < abs(timestamps[ref]):
            ref = i
        else:
            break
    if not mixer or n == 1:
        return [(ref, 1.0)]
    blur = 1.0
    for i in range(1, n):
        if timestamps[i] >= 0.0 and timestamps[i - 1] < 0:
            sample_dur = max(f32(timestamps[i]) - f32(timestamps[i - 1]), f32(vsync))
            if sample_dur > 1.0:
                blur *= float(sample_dur)
            break
    out = []
    for i in range(n):
        rts = f32(timestamps[i])
        if mixer == "oversample":
            end = f32(timestamps[i + 1]) if i + 1 < n else f32(np.inf)
            if rts > vsync or end < 0.0:
                continue
            rts, end = max(rts, f32(0.0)), min(end, f32(vsync))
            w = f32(f32(end - rts) / f32(vsync))
        else:
            # pl_filter_radius_bound: kernel radius * blur (the presets used here are not resizable)
            radius = f32({"mitchell_clamp": 2.0, "linear": 1.0, "hermite": 1.0, "catmull_rom": 2.0}[mixer] * blur)
            if abs(rts) >= radius:
                continue
            w = f32(t0.filter_sample(mixer, float(rts), blur=blur))
        if abs(w) <= 1e-3 and i != ref:
            continue
        out.append((i, float(w)))
    return out or [(ref, 1.0)]


def render_mix(imgs, timestamps, vsync, tgt, mixer="", **kw):
    """pl_render_image_mix (renderer.c:3612-4025): every contributing frame is rendered up to the
    target colour space into an rgba16hf image, the images are linearised, summed with the
    normalised weights, delinearised and sent through the output stage."""
    t0, orc = T.t0(), T.oracle()
    ws = mix_weights(timestamps, vsync, mixer, t0)
    fbos = []
    for i, _ in ws:
        c, _ = render(imgs[i], tgt, _return_color=True, **kw)
        fbos.append(orc.store_round(c, 0))  # cached frame: rgba16hf at the output size
    if len(ws) == 1:
        color = fbos[0]
    else:
        _, dst_csp = t0.csp_infer_map(imgs[0].csp.copy(), tgt.csp)
        dst_trc = int(dst_csp[1])
        dmin, dmax, _ = t0.nominal_luma(dst_csp, T.META["HDR10"], T.HDR["NORM"])
        dluma = luma_coeffs(t0, int(dst_csp[0]))
        wsum = f32(0.0)
        for _, w in ws:
            wsum = f32(wsum + f32(w))
        acc = np.zeros_like(fbos[0])
        for (_, w), fb in zip(ws, fbos):
            lin = orc.linearize(fb, dst_trc, dmin, dmax, dluma)
            acc = orc.mix_accumulate(acc, lin, float(f32(w) / wsum))
        # re-encode to the target transfer: pl_shader_color_map_ex(NULL, linear -> target)
        lin_csp = dst_csp.copy()
        lin_csp[1] = T.TRC["LINEAR"]
        cm = color_map_params(t0, lin_csp, dst_csp, state=False)
        color = orc.color_map(acc, cm["params"]) if cm is not None else acc
    out, _ = render(imgs[0], tgt, _start_color=color, **kw)
    return out


def planar_output(orc, t0, color, tgt, tnorm, depth, dither, dst_trc, frame_index, coord_mode):
    """pass_output_target with several target planes (renderer.c:2796-2961): the encoded image
    goes to an rgba16hf FBO; every plane samples it (1:1 -> nearest; subsampled planes through
    the plane down-scaler, none configured = bilinear pl_shader_sample_direct) over its rounded,
    possibly over-rendered rect, dithers, scales, swizzles and stores the part inside the plane."""
    H, W = color.shape[:2]
    fbo = orc.store_round(color, 0)
    outs = []
    for i, (fmt, comps, (ssx, ssy)) in enumerate(tgt.planes):
        pw, ph = tgt.plane_size(i)
        rrx, rry = f32(1.0 / ssx), f32(1.0 / ssy)
        sx, sy = CHROMA_SHIFT[tgt.chroma_loc] if (ssx > 1 or ssy > 1) else (0.0, 0.0)
        sx, sy = f32(sx), f32(sy)
        x0f, y0f = (f32(0) - sx) * rrx, (f32(0) - sy) * rry
        x1f, y1f = (f32(W) - sx) * rrx, (f32(H) - sy) * rry
        rx0, ry0 = int(np.floor(x0f)), int(np.floor(y0f))
        rx1, ry1 = int(np.ceil(x1f)), int(np.ceil(y1f))
        rect = (float((f32(rx0) - x0f) / rrx), float((f32(ry0) - y0f) / rry),
                float((f32(rx1) - x0f) / rrx), float((f32(ry1) - y0f) / rry))
        nw, nh = rx1 - rx0, ry1 - ry0
        unscaled = ssx == 1 and ssy == 1
        c = orc.sample_direct(fbo, nw, nh, linear=not unscaled, rect=rect, coord_mode=coord_mode)
        if depth and depth < 16 and dither:
            lut = t0.dither_matrix(0 if dither == 1 else 1, 64) if dither in (1, 2) else None
            method = {1: 0, 2: 1, 3: 2, 4: 3}[dither]
            c = orc.dither(c, depth, method, lut, False, frame_index, approx_gamma(dst_trc))
        c = orc.scale(c, float(f32(1.0 / tnorm)), True)
        c = c[max(-ry0, 0):max(-ry0, 0) + min(ry1, ph) - max(ry0, 0),
              max(-rx0, 0):max(-rx0, 0) + min(rx1, pw) - max(rx0, 0)][..., list(comps)]
        bits = 8 if fmt.endswith("8") else 16
        o = orc.store_unorm(np.ascontiguousarray(c), bits)
        outs.append(o.astype(np.uint8) if bits == 8 else o)
    return outs


def approx_gamma(trc):
    g = {0: 1.0, 3: 1.0, 17: 1.0, 10: 1.8, 4: 1.8, 5: 2.0, 7: 2.4, 8: 2.6, 11: 2.6, 9: 2.8,
         2: 2.2, 1: 2.2, 6: 2.2}
    return g.get(trc, 2.0)


def store(orc, color, fmt):
    if fmt == "rgba16hf":
        return orc.store_half(color)
    if fmt == "rgba16":
        return orc.store_unorm(color, 16)
    if fmt == "rgba8":
        return orc.store_unorm(color, 8).astype(np.uint8)
    return color


_keepalive = []


def color_map_params(t0, src, dst, hq=False, tone_fn=None, gamut_fn=None, state=True):
    """Host logic of pl_shader_color_map_ex -> OrcColorMap (or None when src == dst)."""
    src, dst = t0.csp_infer_map(src, dst)
    if np.array_equal(src, dst):
        return None
    tone_fn = tone_fn or "spline"
    gamut_fn = gamut_fn or "perceptual"
    if not state:  # colorspace.c:1746-1753: no state object -> no LUT-based methods
        tone_fn = tone_fn if tone_fn == "clip" else "linear"
        gamut_fn = gamut_fn if gamut_fn == "clip" else "saturation"
    imin, imax, iavg = t0.nominal_luma(src, T.META["ANY"], T.HDR["PQ"])
    omin, omax, _ = t0.nominal_luma(dst, T.META["HDR10"], T.HDR["PQ"])
    tone = t0.tone_map_generate(tone_fn, imin, imax, iavg, omin, omax, hdr_csp=src)
    tpar = tone["params"]
    imin, imax, iavg, omin, omax = (f32(v) for v in tpar[3:8])
    if abs(float(imax) - float(omax)) < 1e-6:
        omax = imax
    if abs(float(imin) - float(omin)) < 1e-6:
        omin = imin
    omax = min(omax, imax)  # no inverse tone mapping
    # regenerate with the adjusted range (infer is idempotent)
    tone = t0.tone_map_generate(tone_fn, imin, imax, iavg, omin, omax, hdr_csp=src)
    need_tone = not tone["noop"]

    gmin, gmax, _ = t0.nominal_luma(dst, T.META["HDR10"], T.HDR["PQ"])
    in_prim, out_prim = src[2:10].copy(), dst[2:10].copy()
    bidir = gamut_fn in ("perceptual", "saturation")
    if bidir and t0.primaries_compatible(in_prim, out_prim):
        out_prim, _ = t0.primaries_clip(out_prim, in_prim)
    g = t0.gamut_map_generate(gamut_fn, in_prim, out_prim, gmin, gmax, generate=False)
    need_gamut = not g["noop"]

    rgb2lms = t0.matrix(2, t0.raw_primaries(int(src[0])))
    lms2rgb = t0.matrix(3, t0.raw_primaries(int(dst[0])))
    if need_gamut and gamut_fn == "saturation":
        # can_fast (colorspace.c:1772-1779): saturation mapping is a pure matrix
        lms2src, dst2lms = t0.matrix(3, in_prim), t0.matrix(2, out_prim)
        lms2rgb = mat_mul(mat_mul(lms2rgb, dst2lms), lms2src)
        need_gamut = False
    smin, smax, _ = t0.nominal_luma(src, T.META["HDR10"], T.HDR["NORM"])
    dmin, dmax, _ = t0.nominal_luma(dst, T.META["HDR10"], T.HDR["NORM"])
    p = T.OrcColorMap()
    p.src_trc, p.dst_trc = int(src[1]), int(dst[1])
    p.src_min, p.src_max, p.dst_min, p.dst_max = smin, smax, dmin, dmax
    p.src_luma = (C.c_float * 3)(*luma_coeffs(t0, int(src[0])))
    p.dst_luma = (C.c_float * 3)(*luma_coeffs(t0, int(dst[0])))
    black_scaled = int(dst[1]) not in (T.TRC["BT_1886"], T.TRC["PQ"], T.TRC["SCRGB"],
                                       T.TRC["V_LOG"], T.TRC["S_LOG1"], T.TRC["S_LOG2"], T.TRC["HLG"])
    p.dst_black_scaled = int(black_scaled)
    p.prelinearized = 0
    p.need_tone_map, p.need_gamut_map = int(need_tone), int(need_gamut)
    if not need_tone and not need_gamut:
        # fast path: plain primaries conversion
        m = mat_mul(lms2rgb, rgb2lms) if int(src[0]) != int(dst[0]) else np.eye(3, dtype=f32).ravel()
        p.lms2rgb = (C.c_float * 9)(*m)
        p.rgb2lms = (C.c_float * 9)(*rgb2lms)
        return dict(params=p)
    p.rgb2lms = (C.c_float * 9)(*rgb2lms)
    p.lms2rgb = (C.c_float * 9)(*lms2rgb)
    keep = []
    if need_tone:
        lut = np.ascontiguousarray(tone["lut"], f32)
        keep.append(lut)
        # can_fast (default params): clip and linear are analytic (colorspace.c:1824-1846)
        p.tone_kind = 1 if tone_fn == "clip" else 2 if tone_fn == "linear" else 0
        p.tone_lut, p.tone_lut_size = T.fp(lut), lut.size
        p.tone_input_min, p.tone_input_max = imin, imax
        p.tone_output_min, p.tone_output_max = omin, omax
        p.tone_exposure = float(tpar[18])  # pl_tone_map_constants.exposure
    if need_gamut:
        full = t0.gamut_map_generate(gamut_fn, in_prim, out_prim, gmin, gmax)["lut"].reshape(-1, 3)
        q = np.zeros((full.shape[0], 4), np.uint16)
        # fill_gamut_lut (src/shaders/colorspace.c:1589-1610)
        q[:, 0] = c_roundf(full[:, 0] * f32(65535))
        q[:, 1] = c_roundf(full[:, 1] * f32(65535) + f32(32767))
        q[:, 2] = c_roundf(full[:, 2] * f32(65535) + f32(32767))
        glut = T.unorm16_to_float(q)
        keep.append(glut)
        p.gamut_lut = T.fp(glut)
        p.lut_I, p.lut_C, p.lut_h = 48, 32, 256
        p.gamut_min_luma, p.gamut_max_luma = gmin, gmax
    _keepalive.append(keep)
    if len(_keepalive) > 8:
        _keepalive.pop(0)
    return dict(params=p, tone=tone, keep=keep)


def c_roundf(a):
    """C roundf (half away from zero) on float32 arrays, result as uint16"""
    a = np.asarray(a, f32)
    return (np.sign(a) * np.floor(np.abs(a) + f32(0.5))).astype(np.int64).astype(np.uint16)


def mat_mul(a, b):
    """pl_matrix3x3_mul in float32, same summation order (src/common.c:171-182)"""
    a, b = a.reshape(3, 3).astype(f32), b.reshape(3, 3).astype(f32)
    o = np.zeros((3, 3), f32)
    for r in range(3):
        for c in range(3):
            o[r, c] = f32(f32(f32(a[r, 0] * b[0, c]) + f32(a[r, 1] * b[1, c])) + f32(a[r, 2] * b[2, c]))
    return o.ravel()


SPECIAL_SYS = ("BT_2020_C", "BT_2100_PQ", "BT_2100_HLG", "XYZ")


def decode_color(orc, t0, rgba, sys, levels, bits):
    """pl_shader_decode_color (colorspace.c:275-470) without colour adjustment: the transform
    of pl_color_repr_decode (scale folded in), then the system's non-linear part."""
    dec = t0.repr_decode(sys, levels, *bits)
    out = orc.apply_matrix(rgba, dec["m"], dec["c"])
    if sys in [T.SYS[k] for k in SPECIAL_SYS if k in T.SYS]:
        out = orc.decode_special(out, sys)
    return out


def encode_color(orc, t0, rgba, sys, levels, bits):
    """pl_shader_encode_color (colorspace.c:472-600)."""
    out = rgba
    if sys in [T.SYS[k] for k in SPECIAL_SYS if k in T.SYS]:
        out = orc.encode_special(out, sys)
    enc = t0.repr_encode(sys, levels, *bits)
    return orc.apply_matrix(out, enc["m"], enc["c"])
