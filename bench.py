#!/usr/bin/env python
"""Benchmark of the hot path: frames/s of pl_render_image on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

A step = one frame through the whole render pass graph (`pl_render_image`). Default workload is
BASELINE.json configs[2] (4K HDR10 PQ -> BT.709 SDR: deband + EWA-Lanczos chroma upscale + HQ
peak detection + tone/gamut mapping + blue-noise dither), the configuration the headline metric
"4K HDR10->SDR EWA-Lanczos frames/sec" is quoted on. Prints ONE JSON line (rank 0).

  value     frames/s with the input planes already resident in HBM (device timed, CUDA events on
            the library's stream, max over ranks). Consecutive frames use different resident
            texture sets (ring > L2).
  e2e       same metric through the public API with pinned HOST buffers: upload of the three
            planes, render, download of the RGBA result, all inside the timed region.
  roofline  dominant kernel: algorithmic bytes per launch / mean launch duration (cudaEvent
            intervals delivered by the library's info_callback during the timed region).
  cpu_baseline / --impl reference: the CPU restatement of the same pass graph
            (tests/pipeline_oracle.py over oracle/libpl_oracle.so, parameters from the compiled
            reference when oracle/_ref is present) on a bounded sample.

Multi-GPU (torchrun, one rank per GPU): frames of a stream are independent -> each rank renders
its own frames (weak scaling); the only exchange is the 3264-byte peak statistics all-gather
(libplacebo_b200/parallel.py), which is part of every timed step when N > 1.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

L2_BYTES = 126e6


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3_4k_hdr10_to_sdr")
    ap.add_argument("--cpu-sample", default="960x540", help="frame size of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------
# CPU arm: the oracle's restatement of the pass graph (test infrastructure used as a baseline)
# --------------------------------------------------------------------------------------------
def cpu_pipeline_fps(wl_name, sample, repeats=1):
    """Times the CPU restatement on one `sample`-sized frame and scales by the pixel ratio to
    workload frames/s. Returns (fps, cores, description)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import pipeline_oracle as P
    import plbtest as T
    from libplacebo_b200 import stream as S

    wl = S.WORKLOADS[wl_name]
    sw, sh = (int(v) for v in sample.split("x"))
    scale = wl["dst"][0] / wl["src"][0]
    planes = S.synth_yuv420p10(sw, sh, seed=0)
    img = P.Frame(planes, S.SYS[wl["src_sys"]], csp=np.array(wl["src_csp"], np.float32))
    depth = wl["dst_depth"]
    tgt = P.Target(int(sw * scale), int(sh * scale), wl["dst_fmt"], sample_depth=depth,
                   color_depth=depth, csp=np.array(wl["dst_csp"], np.float32))
    o = wl["opts"]
    kw = dict(upscaler=wl["upscaler"], deband=bool(o.get("deband")), peak=o.get("peak", 0),
              dither=o.get("dither", 0), cmap_hq=bool(o.get("cmap_hq")))
    if o.get("grain"):
        kw["film_grain"] = (1, 1 if o["grain"] > 1 else 0)
    P.render(img, tgt, **kw)  # warm caches (filter / gamut LUT generation)
    t0 = time.perf_counter()
    for i in range(repeats):
        P.render(img, tgt, frame_index=i, **kw)
    dt = (time.perf_counter() - t0) / repeats
    frac = (sw * sh) / (wl["src"][0] * wl["src"][1])
    fps = frac / dt
    cores = T.oracle().num_threads()
    desc = (f"{repeats} frame(s) of {sw}x{sh} ({frac:.4f} of the workload's pixels), "
            f"{dt:.2f} s each, scaled by pixel count")
    return fps, cores, desc, dt


def run_reference(args, rank):
    if rank != 0:
        return
    steps = max(1, args.steps)
    # bounded: each step is one sample frame; cap the whole run at a few minutes
    fps, cores, desc, dt = cpu_pipeline_fps(args.workload, args.cpu_sample, repeats=1)
    per_step = max(1, int(round(2.0 / max(dt, 1e-3))))   # sample frames per step: about 2 s of CPU work
    budget = 150.0
    n = max(1, min(steps, int(budget / max(per_step * dt, 1e-3))))
    for _ in range(min(args.warmup, 2)):
        cpu_pipeline_fps(args.workload, args.cpu_sample, repeats=per_step)
    fps, cores, desc, dt = cpu_pipeline_fps(args.workload, args.cpu_sample, repeats=n * per_step)
    desc = f"{n} step(s) of {per_step} sample frame(s): " + desc
    from libplacebo_b200 import stream as S
    wl = S.WORKLOADS[args.workload]
    line = {
        "impl": "reference", "metric": "4K HDR10->SDR EWA-Lanczos frames/sec", "value": fps,
        "unit": "frames/s", "n_gpus": args.gpus, "steps": n, "warmup": args.warmup,
        "ms_per_step": per_step * dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "desc": wl["desc"]},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": desc + "; libplacebo has no CPU pixel path (its shaders are "
                         "GLSL), so this is the oracle's C/OpenMP restatement of the same pass "
                         "graph with parameters from the compiled reference library"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# --------------------------------------------------------------------------------------------
# clocks sampler
# --------------------------------------------------------------------------------------------
class Clocks:
    def __init__(self, index):
        self.samples, self.reasons, self.max = [], set(), None
        self.stop = threading.Event()
        self.index = index
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max = float(f[1])
                for n, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:  # noqa: BLE001
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max,
                "reasons": sorted(self.reasons), "samples": len(s)}


# --------------------------------------------------------------------------------------------
# algorithmic bytes of each pass (DESIGN.md "kernels"): every input texel read once, every
# output texel written once
# --------------------------------------------------------------------------------------------
def pass_bytes(desc, wl):
    sw, sh = wl["src"]
    dw, dh = wl["dst"]
    cw, ch = sw // 2, sh // 2
    out_bpp = {"rgba8": 4, "rgba16": 8, "rgba16hf": 8}[wl["dst_fmt"]]
    has = lambda s: s in desc  # noqa: E731
    import re
    m = re.search(r"\[(\d+)x(\d+)\]", desc)
    pw, ph = (int(m.group(1)), int(m.group(2))) if m else (dw, dh)
    if has("AV1 film grain") and not has("polar scaling"):
        # one plane: r16 / r16hf / rg16hf in -> fp16 FBO out (+ luma for chroma grain)
        if (pw, ph) == (cw, ch):
            return cw * ch * 4 + cw * ch * 4 + sw * sh * 2 // 4, "chroma grain"
        return pw * ph * 2 + pw * ph * 2, "luma grain"
    if has("debanding") and not has("polar scaling") and not has("color"):
        if (pw, ph) == (cw, ch):
            # chroma planes U,V (r16) -> rg16hf FBO at chroma size
            return 2 * cw * ch * 2 + cw * ch * 4, "chroma deband"
        return pw * ph * 2 + pw * ph * 2, "luma deband"
    if has("polar scaling") and (has("color decoding") or has("color matrix") or has("linearization")):
        # merged plane pass: luma r16 + chroma (rg16 / rg16hf) -> rgba16hf FBO (or target)
        out = sw * sh * 8
        return sw * sh * 2 + cw * ch * 4 + out, "merged plane pass (chroma EWA + decode)"
    if has("polar scaling"):
        # main scaler: rgba16hf FBO at source size -> output size
        last = not has("peak detection")
        return sw * sh * 8 + dw * dh * (out_bpp if last and not wl["opts"].get("peak") else 8), "main EWA scaler"
    if has("tone / gamut mapping") or has("dithering") or has("delinearization"):
        return dw * dh * 8 + dw * dh * out_bpp, "colour map + output"
    return (sw * sh * 2 + 2 * cw * ch * 2) + dw * dh * out_bpp, "other"


def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch

    from libplacebo_b200 import stream as S

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libplacebo_b200 has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    wl = S.WORKLOADS[args.workload]
    sw, sh = wl["src"]
    dw, dh = wl["dst"]
    out_bpp = {"rgba8": 4, "rgba16": 8, "rgba16hf": 8}[wl["dst_fmt"]]
    in_bytes = sw * sh * 2 + 2 * (sw // 2) * (sh // 2) * 2
    out_bytes = dw * dh * out_bpp
    ring = max(2, int(L2_BYTES * 2 / (in_bytes + out_bytes)) + 1)
    ring = min(ring, 16)
    steps, warmup = args.steps, max(args.warmup, 3)
    # (exactly `steps` frames are timed; the per-pass timers keep the first 256 runs of each pass,
    # so with more steps the pass table averages over the first 256 - warmup of them)

    st = S.Stream(device=local_rank).configure(wl, ring=ring)
    lib_stream = torch.cuda.ExternalStream(st.cuda_stream, device=local_rank)

    # synthetic frames, pinned host staging
    host_in = []
    for r in range(ring):
        y, u, v = S.synth_yuv420p10(sw, sh, seed=rank * 100 + r)
        pin = [torch.from_numpy(p).pin_memory() for p in (y, u, v)]
        host_in.append(pin)
        st.upload(r, *pin)
    host_out = [torch.empty((dh, dw, 4), dtype=torch.uint8 if out_bpp == 4 else torch.uint16).pin_memory()
                for _ in range(2)]
    st.finish()

    use_peak = bool(wl["opts"].get("peak"))
    par = None
    if world > 1 and use_peak:
        from libplacebo_b200 import parallel
        par = parallel.PeakExchange(st, dist, rank, world)

    def frame(i, e2e):
        slot = i % ring
        if e2e:
            st.upload(slot, *host_in[slot])
        if par is not None:
            par.render(slot, slot, i)
        else:
            st.render(slot, slot, frame_index=i % 256)
        if e2e:
            st.download(slot, host_out[i % 2])

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(n, e2e, collect):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        if collect:
            st.collect(True)
        l0 = st.launch_count()
        ev0.record(lib_stream)
        for i in range(n):
            frame(i, e2e)
        if e2e:
            st.join()  # the timed interval ends when the last download has landed
        ev1.record(lib_stream)
        st.finish()
        barrier()
        launches = st.launch_count() - l0
        if collect:
            st.collect(False)
        ms = ev0.elapsed_time(ev1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    for i in range(warmup):
        frame(i, False)
    st.finish()
    with Clocks(local_rank) as clk:
        ms, launches = timed(steps, False, True)
    # deliver the last in-window pass samples (untimed flush frames)
    for i in range(2):
        frame(i, False)
        st.finish()
    passes = st.passes()

    for i in range(3):
        frame(i, True)
    st.finish()
    e2e_steps = max(5, steps // 2)
    ms_e2e, _ = timed(e2e_steps, True, False)

    if rank != 0:
        return
    fps = world * steps / (ms / 1e3)
    fps_e2e = world * e2e_steps / (ms_e2e / 1e3)

    # roofline of the dominant pass
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6560.6))
    plist = []
    for cnt, ns, desc in passes:
        b, label = pass_bytes(desc, wl)
        plist.append({"pass": label, "steps": desc, "launches": cnt, "mean_us": ns / 1e3,
                      "algorithmic_MB": b / 1e6,
                      "GBps": (b / (ns * 1e-9) / 1e9) if ns else None})
    plist.sort(key=lambda p: -(p["mean_us"] * p["launches"]))
    top = plist[0] if plist else None
    roof = None
    # DRAM bytes of the same kernel from the committed `ncu --set full` capture (tools/ncu_traffic.py)
    traffic = None
    try:
        tdb = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tdb.get(args.workload, {}).get(top["pass"] if top else "", {}).get("dram_bytes")
    except Exception:  # noqa: BLE001
        pass
    if top and top["GBps"]:
        roof = {"bound": "hbm", "achieved": round(top["GBps"], 1), "peak": hbm_peak, "unit": "GB/s",
                "frac": round(top["GBps"] / hbm_peak, 4), "traffic": traffic,
                "kernel": top["pass"], "mean_us": round(top["mean_us"], 2),
                "algorithmic_bytes_per_launch": int(top["algorithmic_MB"] * 1e6),
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)",
                "note": "kernel is FP32/FP64-issue bound, not HBM bound; see DESIGN.md"}

    line = {
        "metric": "4K HDR10->SDR EWA-Lanczos frames/sec", "value": round(fps, 2), "unit": "frames/s",
        "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": round(ms / steps, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "desc": wl["desc"], "src": [sw, sh], "dst": [dw, dh],
                   "dst_fmt": wl["dst_fmt"], "frames_per_gpu_per_step": 1,
                   "l2": f"ring of {ring} resident input/output texture sets "
                         f"({ring * (in_bytes + out_bytes) / 1e6:.0f} MB > 126 MB L2)",
                   "parallelism": f"frame-parallel x{world}"},
        "roofline": roof,
        "passes": plist,
        "e2e": {"value": round(fps_e2e, 2), "unit": "frames/s", "h2d_bytes_per_step": in_bytes,
                "d2h_bytes_per_step": out_bytes, "steps": e2e_steps,
                "ms_per_step": round(ms_e2e / e2e_steps, 4)},
        "gpu_launches": launches,
        "clocks": clk.summary(),
    }
    if world == 1 and not args.no_cpu_baseline:
        try:
            _, _, _, dt1 = cpu_pipeline_fps(args.workload, args.cpu_sample, repeats=1)
            reps = max(1, min(400, int(round(12.0 / max(dt1, 1e-3)))))   # about 12 s of CPU work
            cfps, cores, desc, _ = cpu_pipeline_fps(args.workload, args.cpu_sample, repeats=reps)
            line["cpu_baseline"] = {"value": round(cfps, 4), "unit": "frames/s", "cores": cores,
                                    "kind": "port", "sample": desc}
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": 0, "kind": "port",
                                    "sample": f"failed: {e}"}
    emit(line)


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the process's real stdout; everything else printed during the
    run (NCCL's version banner, library logs) has been redirected to stderr."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    global _REAL_STDOUT
    args = parse()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # native libraries (NCCL) print to fd 1
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world == 1 and args.gpus > 1:
        emit({"error": f"--gpus {args.gpus} needs torchrun (one rank per GPU)"})
        sys.exit(2)
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
