"""numpy-level wrapper over tests/c/libplb_shim.so (which drives the public C ABI)."""
import ctypes as C

import numpy as np

import plbtest as T

FMT = {
    "r8": (np.uint8, 1), "rg8": (np.uint8, 2), "rgba8": (np.uint8, 4),
    "r16": (np.uint16, 1), "rg16": (np.uint16, 2), "rgba16": (np.uint16, 4),
    "r16hf": (np.float16, 1), "rg16hf": (np.float16, 2), "rgba16hf": (np.float16, 4),
    "r32f": (np.float32, 1), "rg32f": (np.float32, 2), "rgba32f": (np.float32, 4),
}

SRC_KIND = dict(direct=0, nearest=1, bilinear=2, deband=3, polar=4, ortho=5, bicubic=6, hermite=7,
                gaussian=8, oversample=9)
OP = dict(decode=1, encode=2, linearize=3, delinearize=4, color_map=5, dither=6, sigmoidize=7,
          unsigmoidize=8, detect_peak=9)


class Gpu:
    def __init__(self, literal_coords=False):
        self.lib = T.shim()
        self.lib.shim_create.restype = C.c_void_p
        self.lib.shim_tex_create.restype = C.c_void_p
        self.lib.shim_launch_count.restype = C.c_ulonglong
        self.ctx = self.lib.shim_create(int(literal_coords))
        if not self.ctx:
            raise RuntimeError("pl_cuda_create failed (no GPU?) — there is no CPU fallback")

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def close(self):
        if self.ctx:
            self.lib.shim_destroy(C.c_void_p(self.ctx))
            self.ctx = None

    def launch_count(self):
        return int(self.lib.shim_launch_count(C.c_void_p(self.ctx)))

    def set_frame_index(self, i):
        self.lib.shim_set_frame_index(C.c_void_p(self.ctx), int(i))

    def tex(self, fmt, w, h, data=None):
        dt, comps = FMT[fmt]
        ptr = None
        if data is not None:
            data = np.ascontiguousarray(data, dt).reshape(h, w, comps)
            ptr = data.ctypes.data_as(C.c_void_p)
        t = self.lib.shim_tex_create(C.c_void_p(self.ctx), fmt.encode(), w, h, ptr)
        if not t:
            raise RuntimeError(f"pl_tex_create({fmt}, {w}x{h}) failed")
        return t

    def free(self, *texs):
        for t in texs:
            self.lib.shim_tex_destroy(C.c_void_p(self.ctx), C.c_void_p(t))

    def download(self, tex, fmt, w, h):
        dt, comps = FMT[fmt]
        out = np.zeros((h, w, comps), dt)
        r = self.lib.shim_tex_download(C.c_void_p(self.ctx), C.c_void_p(tex),
                                       out.ctypes.data_as(C.c_void_p))
        if r:
            raise RuntimeError("pl_tex_download failed")
        return out

    def run(self, src_tex, dst_tex, kind, sargs, filter_name=None, ops=(), oargs=(), want_peak=False):
        sa = np.zeros(16, T.f32)
        sa[:len(sargs)] = sargs
        opa = np.array(list(ops) + [0], np.int32)
        oa = np.ascontiguousarray(np.concatenate([np.atleast_1d(np.asarray(a, T.f32)) for a in oargs])
                                  if len(oargs) else np.zeros(1), T.f32)
        peak = np.zeros(2, T.f32)
        r = self.lib.shim_run(C.c_void_p(self.ctx), C.c_void_p(src_tex), C.c_void_p(dst_tex),
                              SRC_KIND[kind], T.fp(sa), (filter_name or "").encode(), T.ip(opa),
                              T.fp(oa), T.fp(peak) if want_peak else None)
        if r:
            raise RuntimeError(f"shim_run({kind}) failed with {r}")
        return peak

    # ---- conveniences -----------------------------------------------------------------
    def polar(self, src, src_fmt, dw, dh, filt, comps=4, out_fmt="rgba32f", rect=(0, 0, 0, 0),
              scale=0.0, antiring=-1.0, ops=(), oargs=()):
        h, w = src.shape[:2]
        s = self.tex(src_fmt, w, h, src)
        d = self.tex(out_fmt, dw, dh)
        try:
            self.run(s, d, "polar", list(rect) + [dw, dh, comps, scale, antiring], filt, ops, oargs)
            return self.download(d, out_fmt, dw, dh)
        finally:
            self.free(s, d)

    def simple(self, kind, src, src_fmt, dw, dh, out_fmt="rgba32f", sargs_extra=(), rect=(0, 0, 0, 0),
               comps=4, scale=0.0, ops=(), oargs=(), filt=None, want_peak=False, address=0):
        h, w = src.shape[:2]
        s = self.tex(src_fmt, w, h, src)
        d = self.tex(out_fmt, dw, dh)
        try:
            sargs = list(rect) + [dw, dh, comps, scale] + list(sargs_extra)
            sargs += [0.0] * (15 - len(sargs)) + [float(address)]
            peak = self.run(s, d, kind, sargs, filt, ops, oargs, want_peak)
            out = self.download(d, out_fmt, dw, dh)
            return (out, peak) if want_peak else out
        finally:
            self.free(s, d)
