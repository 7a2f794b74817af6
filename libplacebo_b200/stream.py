"""ctypes front-end to libplb_pyglue.so: render planar YUV frames with pl_render_image.

This is plumbing for bench.py and the multi-GPU driver (parallel.py): the pixels are produced
by the CUDA library behind the C ABI; nothing here computes on the CPU. Workload presets mirror
BASELINE.json's configs.
"""
import ctypes as C
import os

import numpy as np

from . import load, _HERE

GLUE_PATH = os.path.join(_HERE, "libplb_pyglue.so")
PEAK_BUF_SIZE = 3264  # PL_PEAK_BUF_SIZE

# enum values of include/libplacebo_b200.h (identical to libplacebo's colorspace.h)
SYS = dict(UNKNOWN=0, BT_601=1, BT_709=2, SMPTE_240M=3, BT_2020_NC=4, BT_2020_C=5, BT_2100_PQ=6,
           BT_2100_HLG=7, DOLBYVISION=8, YCGCO=9, YCGCO_RE=10, YCGCO_RO=11, RGB=12, XYZ=13)
LEVELS = dict(UNKNOWN=0, LIMITED=1, FULL=2)
PRIM = dict(UNKNOWN=0, BT_601_525=1, BT_601_625=2, BT_709=3, BT_470M=4, EBU_3213=5, BT_2020=6)
TRC = dict(UNKNOWN=0, BT_1886=1, SRGB=2, LINEAR=3, PQ=12, HLG=13)
CHROMA = dict(left=1, center=2, top_left=3)
_OPT_ORDER = ["src_w", "src_h", "chroma_w", "chroma_h", "dst_w", "dst_h", "src_sys", "src_levels",
              "src_sample_depth", "src_color_depth", "src_bit_shift", "dst_sys", "dst_levels",
              "dst_sample_depth", "dst_color_depth", "chroma_loc", "deband", "peak",
              "allow_delayed", "dither", "cmap_hq", "ring", "tone_fn", "gamut_fn", "grain", "dovi", "errdiff", "contrast"]
# pl_error_diffusion_kernels[] order (dither.c)
ERRDIFF = ["simple", "false-fs", "sierra-lite", "floyd-steinberg", "atkinson", "jarvis-judice-ninke",
           "stucki", "burkes", "sierra-2", "sierra-3"]
_FMT_DTYPE = {"rgba8": (np.uint8, 4), "rgba16": (np.uint16, 4), "rgba16hf": (np.uint16, 4),
              "rgba32f": (np.float32, 4)}


def csp(prim=0, trc=0, min_luma=0.0, max_luma=0.0):
    """Flat pl_color_space: prim, trc, hdr.prim[8], min_luma, max_luma, max_cll, max_fall,
    scene_max[3], scene_avg, max_pq_y, avg_pq_y."""
    a = np.zeros(20, np.float32)
    a[0], a[1], a[10], a[11] = prim, trc, min_luma, max_luma
    return a


# Workloads (BASELINE.json `configs`): C2 = configs[1], C3 = configs[2]
WORKLOADS = {
    "c2_1080p_to_4k_ewa_fp16": dict(
        desc="1080p YUV420P10 BT.709 -> 4K rgba16hf, EWA-Lanczos main + chroma upscale",
        src=(1920, 1080), dst=(3840, 2160), dst_fmt="rgba16hf",
        src_sys="BT_709", src_csp=csp(PRIM["BT_709"], TRC["BT_1886"]),
        dst_csp=csp(PRIM["BT_709"], TRC["BT_1886"]), dst_depth=0,
        upscaler="ewa_lanczos", opts=dict()),
    "c3_4k_hdr10_to_sdr": dict(
        desc="4K YUV420P10 HDR10 PQ BT.2020 -> 4K rgba8 BT.709: deband + EWA-Lanczos chroma "
             "upscale + HQ peak detection + spline tone map + perceptual gamut map + blue-noise "
             "dither",
        src=(3840, 2160), dst=(3840, 2160), dst_fmt="rgba8",
        src_sys="BT_2020_NC", src_csp=csp(PRIM["BT_2020"], TRC["PQ"], max_luma=1000),
        dst_csp=csp(PRIM["BT_709"], TRC["BT_1886"]), dst_depth=8,
        upscaler="ewa_lanczos", opts=dict(deband=1, peak=2, dither=1, cmap_hq=1)),
    "c3_4k_hdr10_to_sdr_hq_preset": dict(
        desc="config C3 with pl_color_map_high_quality_params unchanged: contrast recovery 0.3 (PQ "
             "intensity feature map, bicubic low-pass by 3.5, bicubic detail fetch in the tone map)",
        src=(3840, 2160), dst=(3840, 2160), dst_fmt="rgba8",
        src_sys="BT_2020_NC", src_csp=csp(PRIM["BT_2020"], TRC["PQ"], max_luma=1000),
        dst_csp=csp(PRIM["BT_709"], TRC["BT_1886"]), dst_depth=8,
        upscaler="ewa_lanczos", opts=dict(deband=1, peak=2, dither=1, cmap_hq=1, contrast=1)),
    "c4_4k_hdr10_to_sdr_av1_grain": dict(
        desc="config C3 + AV1 film grain synthesis (libplacebo test-fixture parameters, block "
             "overlap, new seed every frame) on every plane before scaling",
        src=(3840, 2160), dst=(3840, 2160), dst_fmt="rgba8",
        src_sys="BT_2020_NC", src_csp=csp(PRIM["BT_2020"], TRC["PQ"], max_luma=1000),
        dst_csp=csp(PRIM["BT_709"], TRC["BT_1886"]), dst_depth=8,
        upscaler="ewa_lanczos", opts=dict(deband=1, peak=2, dither=1, cmap_hq=1, grain=2)),
    "c5_8k_dovi_errdiff": dict(
        desc="8K YUV420P10 Dolby Vision (8-piece polynomial luma + order-3 MMR chroma reshaping, RPU "
             "matrices) PQ BT.2020 -> 8K rgba8 BT.709: EWA-Lanczos chroma upscale + spline tone map + "
             "perceptual gamut map + sierra-lite error diffusion to 8 bits",
        src=(7680, 4320), dst=(7680, 4320), dst_fmt="rgba8",
        src_sys="BT_2020_NC", src_levels="FULL", src_csp=csp(PRIM["BT_2020"], TRC["PQ"], max_luma=1000),
        dst_csp=csp(PRIM["BT_709"], TRC["BT_1886"]), dst_depth=8,
        upscaler="ewa_lanczos", opts=dict(dither=1, cmap_hq=1, dovi=1, errdiff=ERRDIFF.index("sierra-lite") + 1)),
    # profiling workloads for the kernel families the BASELINE configs do not reach (generic-ratio
    # EWA, separable scalers); not benchmark configurations
    "p1_1080p_to_1440p_ewa_generic": dict(
        desc="1080p YUV420P10 BT.709 -> 2560x1440 rgba16hf, EWA-Lanczos at ratio 4/3 (generic polar kernel)",
        src=(1920, 1080), dst=(2560, 1440), dst_fmt="rgba16hf",
        src_sys="BT_709", src_csp=csp(PRIM["BT_709"], TRC["BT_1886"]),
        dst_csp=csp(PRIM["BT_709"], TRC["BT_1886"]), dst_depth=0,
        upscaler="ewa_lanczos", opts=dict()),
    "p2_1080p_to_1440p_lanczos_ortho": dict(
        desc="1080p YUV420P10 BT.709 -> 2560x1440 rgba16hf, separable Lanczos (two ortho passes), "
             "separable chroma upscale",
        src=(1920, 1080), dst=(2560, 1440), dst_fmt="rgba16hf",
        src_sys="BT_709", src_csp=csp(PRIM["BT_709"], TRC["BT_1886"]),
        dst_csp=csp(PRIM["BT_709"], TRC["BT_1886"]), dst_depth=0,
        upscaler="lanczos", opts=dict()),
    # reduced-size C3 for the CPU-only contract tests of bench.py (not a benchmark configuration)
    "test_c3_480x270": dict(
        desc="config C3 at 480x270 (bench.py contract tests)",
        src=(480, 270), dst=(480, 270), dst_fmt="rgba8",
        src_sys="BT_2020_NC", src_csp=csp(PRIM["BT_2020"], TRC["PQ"], max_luma=1000),
        dst_csp=csp(PRIM["BT_709"], TRC["BT_1886"]), dst_depth=8,
        upscaler="ewa_lanczos", opts=dict(deband=1, peak=2, dither=1, cmap_hq=1)),
}


def synth_yuv420p10(w, h, seed=0):
    """Synthetic 10-bit limited-range frame (uint16 code values): smooth gradients, concentric
    rings and mild noise, so that debanding, the peak histogram and the tone curve all have
    realistic work to do."""
    rng = np.random.default_rng(seed)

    def plane(pw, ph, lo, hi, phase):
        y, x = np.mgrid[0:ph, 0:pw].astype(np.float32)
        cx, cy = pw * (0.35 + 0.1 * phase), ph * (0.55 - 0.1 * phase)
        r = np.hypot(x - cx, y - cy) / max(pw, ph)
        v = 0.5 + 0.35 * np.cos(r * (40.0 + 10 * phase)) * np.exp(-2.0 * r) + 0.15 * (x / pw - 0.5)
        v = lo + np.clip(v, 0, 1) * (hi - lo) + rng.integers(-2, 3, (ph, pw))
        return np.clip(np.rint(v), 0, 1023).astype(np.uint16)

    return (plane(w, h, 64, 940, 0.0 + 0.01 * seed), plane(w // 2, h // 2, 64, 960, 1.0),
            plane(w // 2, h // 2, 64, 960, 2.0))


class Stream:
    """One renderer with a fixed configuration on one GPU."""

    def __init__(self, device=0, literal_coords=False):
        load()  # fails loudly when the CUDA library is missing
        if not os.path.exists(GLUE_PATH):
            raise RuntimeError(f"{GLUE_PATH} is missing: run __graft_entry__.build()")
        self.lib = C.CDLL(GLUE_PATH, mode=C.RTLD_LOCAL)
        L = self.lib
        L.plg_create.restype = C.c_void_p
        L.plg_cuda_stream.restype = C.c_void_p
        L.plg_launch_count.restype = C.c_uint64
        for f in ("plg_destroy", "plg_finish", "plg_reset", "plg_join"):
            getattr(L, f).restype = None
        self.h = L.plg_create(device, int(literal_coords))
        if not self.h:
            raise RuntimeError("pl_cuda_create failed: no usable CUDA device (there is no CPU "
                               "fallback)")
        self.h = C.c_void_p(self.h)
        self.cfg = None

    def close(self):
        if self.h:
            self.lib.plg_destroy(self.h)
            self.h = None

    def configure(self, wl, ring=1, **over):
        """`wl`: a WORKLOADS entry (or equivalent dict)."""
        o = dict(src_w=wl["src"][0], src_h=wl["src"][1], chroma_w=wl["src"][0] // 2,
                 chroma_h=wl["src"][1] // 2, dst_w=wl["dst"][0], dst_h=wl["dst"][1],
                 src_sys=SYS[wl["src_sys"]], src_levels=LEVELS[wl.get("src_levels", "LIMITED")],
                 src_sample_depth=16,
                 src_color_depth=10, src_bit_shift=0, dst_sys=SYS["RGB"], dst_levels=LEVELS["FULL"],
                 dst_sample_depth=wl["dst_depth"], dst_color_depth=wl["dst_depth"],
                 chroma_loc=CHROMA["left"], deband=0, peak=0, allow_delayed=0, dither=0,
                 cmap_hq=0, ring=ring, tone_fn=0, gamut_fn=0, grain=0, dovi=0, errdiff=0, contrast=0)
        o.update(wl.get("opts", {}))
        o.update(over)
        arr = np.array([o[k] for k in _OPT_ORDER], np.int32)
        s, d = np.ascontiguousarray(wl["src_csp"], np.float32), np.ascontiguousarray(wl["dst_csp"], np.float32)
        up = wl.get("upscaler") or ""
        r = self.lib.plg_configure(self.h, arr.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p),
                                   d.ctypes.data_as(C.c_void_p), b"r16", wl["dst_fmt"].encode(),
                                   up.encode(), up.encode())
        if r:
            raise RuntimeError(f"plg_configure failed: {r}")
        self.cfg = dict(o, dst_fmt=wl["dst_fmt"])
        return self

    # ---- per-frame calls (asynchronous on the library's stream) --------------------------
    def upload(self, slot, y, u, v):
        """y, u, v: C-contiguous uint16 host arrays (numpy or pinned torch .numpy())."""
        r = self.lib.plg_upload(self.h, slot, C.c_void_p(_ptr(y)), C.c_void_p(_ptr(u)), C.c_void_p(_ptr(v)))
        if r:
            raise RuntimeError("pl_tex_upload failed")

    def render(self, slot=0, out_slot=0, frame_index=-1):
        r = self.lib.plg_render(self.h, slot, out_slot, frame_index)
        if r:
            raise RuntimeError(f"pl_render_image failed or degraded (errors={r})")

    def download(self, out_slot, host):
        if self.lib.plg_download(self.h, out_slot, C.c_void_p(_ptr(host))):
            raise RuntimeError("pl_tex_download failed")

    def out_array(self):
        dt, comps = _FMT_DTYPE[self.cfg["dst_fmt"]]
        return np.zeros((self.cfg["dst_h"], self.cfg["dst_w"], comps), dt)

    def finish(self):
        self.lib.plg_finish(self.h)

    def join(self):
        """Main stream waits for all transfers enqueued so far (see pl_cuda_join)."""
        self.lib.plg_join(self.h)

    def reset(self):
        self.lib.plg_reset(self.h)

    @property
    def cuda_stream(self):
        return self.lib.plg_cuda_stream(self.h)

    def launch_count(self):
        return int(self.lib.plg_launch_count(self.h))

    # ---- frame-parallel peak statistics ---------------------------------------------------
    def measure_peak(self, slot, frame_index=-1):
        """Runs the source pass of the frame in `slot` (kept for the next render of the same
        slot) and returns its raw peak statistics. `frame_index` must be the one the following
        render() uses (it seeds the debanding noise and the grain)."""
        buf = np.zeros(PEAK_BUF_SIZE // 4, np.uint32)
        if self.lib.plg_measure_peak(self.h, slot, frame_index, C.c_void_p(_ptr(buf))):
            raise RuntimeError("pl_renderer_measure_peak failed")
        return buf

    def measure_peak_dev(self, slot, frame_index, dev_ptr):
        """Asynchronous measure_peak: the statistics are copied to device memory at `dev_ptr`
        (PEAK_BUF_SIZE bytes) on the library's stream; nothing waits on the host."""
        if self.lib.plg_measure_peak_dev(self.h, slot, frame_index, C.c_void_p(dev_ptr)):
            raise RuntimeError("pl_renderer_measure_peak_dev failed")

    def small_copy(self, dst_ptr, src_ptr, nbytes, cuda_stream=0):
        """pl_cuda_small_copy: kernel copy (device / pinned host memory) on `cuda_stream`."""
        if self.lib.plg_small_copy(self.h, C.c_void_p(dst_ptr), C.c_void_p(src_ptr), C.c_size_t(nbytes),
                                   C.c_void_p(cuda_stream)):
            raise RuntimeError("pl_cuda_small_copy failed")

    def peak_update(self, state, buf):
        st = np.array(state, np.float32)
        used = self.lib.plg_peak_update(self.h, C.c_void_p(_ptr(st)), C.c_void_p(_ptr(buf)), None)
        return (float(st[0]), float(st[1])), bool(used)

    def set_peak(self, state):
        st = np.array(state if state else (0, 0), np.float32)
        self.lib.plg_set_peak(self.h, C.c_void_p(_ptr(st)))

    def get_peak(self):
        st = np.zeros(2, np.float32)
        return None if self.lib.plg_get_peak(self.h, C.c_void_p(_ptr(st))) else (float(st[0]), float(st[1]))

    # ---- pass statistics --------------------------------------------------------------------
    def collect(self, on):
        self.lib.plg_collect(self.h, int(on))

    def passes(self):
        """[(count, mean_ns, description)] for the last collection window."""
        buf = C.create_string_buffer(16384)
        self.lib.plg_pass_list(self.h, buf, C.c_size_t(len(buf)))
        out = []
        for line in buf.value.decode().splitlines():
            c, ns, desc = line.split("\t", 2)
            out.append((int(c), int(ns), desc))
        return out


def _ptr(a):
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data
