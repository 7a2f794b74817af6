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
  extra_workloads  the same measurement (shorter) for configs[1] (C2, the north-star roofline
            configuration: 1080p -> 4K EWA-Lanczos fp16) and configs[3] (C4: C3 + AV1 grain),
            each with its own `roofline` — on N ranks when launched under torchrun.
  parity    one frame of the timed configuration rendered from a fresh renderer and compared
            with the CPU oracle (max difference in LSB of the target); rank 0, N = 1.
  cpu_baseline / --impl reference: the CPU restatement of the same pass graph
            (tests/pipeline_oracle.py over oracle/libpl_oracle.so, parameters from the compiled
            reference when oracle/_ref is present) on whole frames of the workload, OpenMP
            threads pinned to the physical cores; value = 1 / median frame time.

Multi-GPU (torchrun, one rank per GPU): frames of a stream are independent -> each rank renders
its own frames (weak scaling); the only exchange is the peak statistics all-gather
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
METRIC = "4K HDR10->SDR EWA-Lanczos frames/sec"
EXTRA = ["c2_1080p_to_4k_ewa_fp16", "c4_4k_hdr10_to_sdr_av1_grain", "c3_4k_hdr10_to_sdr_hq_preset"]


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3_4k_hdr10_to_sdr")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip extra_workloads and parity")
    ap.add_argument("--c5-lanes", type=int, default=8, help="frames in flight per GPU for config C5 (0 = skip)")
    return ap.parse_args()


def physical_cores():
    """Physical cores this process may run on (OpenMP threads of the CPU arm are pinned to them)."""
    try:
        allowed = os.sched_getaffinity(0)
        seen = set()
        for cpu in allowed:
            base = f"/sys/devices/system/cpu/cpu{cpu}/topology"
            with open(base + "/physical_package_id") as f:
                pkg = f.read().strip()
            with open(base + "/core_id") as f:
                core = f.read().strip()
            seen.add((pkg, core))
        return max(1, len(seen))
    except Exception:  # noqa: BLE001
        return max(1, (os.cpu_count() or 2) // 2)


def pin_openmp():
    """The CPU arm uses every physical core, whatever the launcher exported (torchrun sets
    OMP_NUM_THREADS=1). Must run before the oracle library (libgomp) is loaded."""
    n = physical_cores()
    os.environ["OMP_NUM_THREADS"] = str(n)
    os.environ["OMP_PROC_BIND"] = "spread"
    os.environ["OMP_PLACES"] = "cores"
    os.environ["OMP_DYNAMIC"] = "false"
    return n


def numa_pin(local_rank):
    """Run this rank on the CPUs local to its GPU (pinned staging memory is then allocated on the
    GPU's NUMA node: host->device copies do not cross the socket interconnect). Returns the
    previous affinity, or None."""
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i",
                              str(local_rank)], capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if bus.count(":") == 2 and len(bus.split(":")[0]) == 8:
            bus = bus[4:]  # 00000000:1b:00.0 -> 0000:1b:00.0
        with open(f"/sys/bus/pci/devices/{bus}/local_cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        prev = os.sched_getaffinity(0)
        cpus &= prev
        if cpus:
            os.sched_setaffinity(0, cpus)
            return prev
    except Exception:  # noqa: BLE001
        pass
    return None


def workload_config(wl_name, world):
    """The `config` object: identical in both arms."""
    from libplacebo_b200 import stream as S
    wl = S.WORKLOADS[wl_name]
    return {"workload": wl_name, "desc": wl["desc"], "src": list(wl["src"]), "dst": list(wl["dst"]),
            "dst_fmt": wl["dst_fmt"], "frames_per_gpu_per_step": 1,
            "parallelism": f"frame-parallel x{world}"}


# --------------------------------------------------------------------------------------------
# CPU arm: the oracle's restatement of the pass graph (test infrastructure used as a baseline)
# --------------------------------------------------------------------------------------------
def oracle_render_args(wl_name, seed=0):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import pipeline_oracle as P
    from libplacebo_b200 import stream as S

    wl = S.WORKLOADS[wl_name]
    sw, sh = wl["src"]
    planes = S.synth_yuv420p10(sw, sh, seed=seed)
    img = P.Frame(planes, S.SYS[wl["src_sys"]], csp=np.array(wl["src_csp"], np.float32))
    depth = wl["dst_depth"]
    tgt = P.Target(wl["dst"][0], wl["dst"][1], wl["dst_fmt"], sample_depth=depth,
                   color_depth=depth, csp=np.array(wl["dst_csp"], np.float32))
    o = wl["opts"]
    kw = dict(upscaler=wl["upscaler"], deband=bool(o.get("deband")), peak=o.get("peak", 0),
              dither=o.get("dither", 0), cmap_hq=bool(o.get("cmap_hq")))
    if o.get("contrast"):
        kw["contrast_recovery"] = 0.3
    return P, img, tgt, kw, bool(o.get("grain"))


def cpu_frames(wl_name, n, warm=1):
    """Times n whole frames of the workload on the CPU restatement. Returns the list of frame
    times (s), the thread count and a description."""
    P, img, tgt, kw, grain = oracle_render_args(wl_name)
    import plbtest  # tests/ is on sys.path after oracle_render_args
    times = []
    for i in range(warm + n):
        k = dict(kw, frame_index=i)
        if grain:
            k["film_grain"] = (i + 1, 1)
        t0 = time.perf_counter()
        P.render(img, tgt, **k)
        dt = time.perf_counter() - t0
        if i >= warm:
            times.append(dt)
    return times, plbtest.oracle().num_threads()


def cpu_arm_line(wl_name, times, cores):
    s = sorted(times)
    med = s[len(s) // 2]
    return {"value": round(1.0 / med, 4), "unit": "frames/s", "cores": cores, "kind": "port",
            "frame_s_median": round(med, 4), "frame_s_min": round(s[0], 4), "frame_s_max": round(s[-1], 4),
            "sample": f"{len(s)} whole frame(s) of {wl_name} (no scaling), median frame time; OpenMP over "
                      f"rows, {cores} threads pinned to physical cores (OMP_PROC_BIND=spread, "
                      "OMP_PLACES=cores); libplacebo has no CPU pixel path (its shaders are GLSL), "
                      "so this is the oracle's C/OpenMP restatement of the same pass graph with "
                      "parameters from the compiled reference library"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    steps = max(1, args.steps)
    # bounded: the whole run ends within a few minutes whatever K is
    t1, cores = cpu_frames(args.workload, 1, warm=1)
    n = max(5, min(steps, int(150.0 / max(t1[0], 1e-3))))
    times, cores = cpu_frames(args.workload, n, warm=min(max(args.warmup, 0), 2))
    cb = cpu_arm_line(args.workload, times, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": cb["value"],
        "unit": "frames/s", "n_gpus": args.gpus, "steps": n, "warmup": args.warmup,
        "ms_per_step": round(cb["frame_s_median"] * 1e3, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, world),
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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
    if has("feature extraction"):
        return dw * dh * 8 + dw * dh * 2, "feature extraction"
    if has("ortho") or has("separable"):
        if wl["opts"].get("contrast"):
            return pw * ph * 2 * 2, "feature low-pass"
        return sw * sh * 8 + pw * ph * 8, "separable scaler pass"
    if has("tone / gamut mapping") or has("dithering") or has("delinearization"):
        return dw * dh * 8 + dw * dh * out_bpp, "colour map + output"
    return (sw * sh * 2 + 2 * cw * ch * 2) + dw * dh * out_bpp, "other"


class Ctx:
    pass


def measure(cx, wl_name, steps, warmup, with_e2e=True):
    """One workload on this rank's GPU (all ranks take part). Returns a dict on every rank."""
    import torch

    from libplacebo_b200 import stream as S
    torch, dist, rank, world, local_rank = cx.torch, cx.dist, cx.rank, cx.world, cx.local_rank

    wl = S.WORKLOADS[wl_name]
    sw, sh = wl["src"]
    dw, dh = wl["dst"]
    out_bpp = {"rgba8": 4, "rgba16": 8, "rgba16hf": 8}[wl["dst_fmt"]]
    in_bytes = sw * sh * 2 + 2 * (sw // 2) * (sh // 2) * 2
    out_bytes = dw * dh * out_bpp
    ring = max(2, int(L2_BYTES * 2 / (in_bytes + out_bytes)) + 1)
    ring = min(ring, 16)
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
    e2e_mode = [False]
    if world > 1 and use_peak:
        from libplacebo_b200 import parallel
        # statistics exchange pipelined one frame ahead (side-stream NCCL all-gather): the
        # frame's upload (end-to-end mode) happens right before it is measured
        par = parallel.PeakPipeline(st, dist, rank, world, ring,
                                    before_measure=lambda j: st.upload(j % ring, *host_in[j % ring])
                                    if e2e_mode[0] else None)

    uploaded = [-1]  # end-to-end mode: index of the newest frame whose upload has been enqueued

    def frame(i, e2e, last=False):
        slot = i % ring
        if par is not None:
            e2e_mode[0] = e2e
            par.render(i, last=last)
        else:
            if e2e:
                # Like a player, upload the NEXT frame before rendering this one: pl_render_image
                # waits on the host for this frame's peak statistics (same-frame peak detection),
                # and the copy engine should not idle meanwhile. One upload per step.
                if uploaded[0] < i:
                    st.upload(slot, *host_in[slot])
                    uploaded[0] = i
                if not last and uploaded[0] < i + 1:
                    st.upload((i + 1) % ring, *host_in[(i + 1) % ring])
                    uploaded[0] = i + 1
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
        uploaded[0] = -1
        ev0.record(lib_stream)
        for i in range(n):
            frame(i, e2e, last=i == n - 1)
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
        frame(i, False, last=i == warmup - 1)
    st.finish()
    with Clocks(local_rank) as clk:
        ms, launches = timed(steps, False, True)
    # deliver the last in-window pass samples (untimed flush frames)
    for i in range(2):
        frame(i, False, last=True)
        st.finish()
    passes = st.passes()

    res = {"wl": wl, "ring": ring, "in_bytes": in_bytes, "out_bytes": out_bytes, "ms": ms,
           "launches": launches, "passes": passes, "clocks": clk.summary(), "steps": steps}
    if with_e2e:
        uploaded[0] = -1
        for i in range(3):
            frame(i, True, last=i == 2)
        st.finish()
        e2e_steps = max(5, steps // 2)
        ms_e2e, _ = timed(e2e_steps, True, False)
        res["ms_e2e"], res["e2e_steps"] = ms_e2e, e2e_steps

    # parity of one frame of this configuration (fresh renderer: no peak history), rank 0 only
    res["parity_frame"] = None
    if cx.want_parity and rank == 0 and world == 1:
        st.finish()
        st.reset()
        st.render(0, 0, frame_index=0)
        out = st.out_array()
        st.download(0, out)
        st.join()
        st.finish()
        res["parity_frame"] = out
    st.close()
    return res


def measure_lanes(cx, wl_name, lanes, frames_per_lane):
    """Config C5: error diffusion is a serial wavefront (one 8-CTA thread-block cluster per frame,
    w + (h-1) * shift barrier steps), so a single frame cannot fill the GPU; throughput comes from frames in
    flight. `lanes` independent pl_gpu + pl_renderer pairs on this rank's device (one CUDA stream
    each) render different frames concurrently: the wavefront passes of some frames overlap
    the wide passes of others. Device-timed from a common start event to the last lane's end."""
    from libplacebo_b200 import stream as S
    torch, rank, local_rank = cx.torch, cx.rank, cx.local_rank
    wl = S.WORKLOADS[wl_name]
    sw, sh = wl["src"]
    y, u, v = S.synth_yuv420p10(sw, sh, seed=rank)
    pin = [torch.from_numpy(p).pin_memory() for p in (y, u, v)]
    sts = []
    try:
        for _ in range(lanes):
            st = S.Stream(device=local_rank).configure(wl, ring=1)
            st.upload(0, *pin)
            sts.append(st)
        for st in sts:
            st.finish()
        ext = [torch.cuda.ExternalStream(st.cuda_stream, device=local_rank) for st in sts]

        def run(nl, nf):
            for st in sts[:nl]:   # warm-up: allocates the FBOs, builds the LUTs
                st.render(0, 0, frame_index=0)
            for st in sts[:nl]:
                st.finish()
            if cx.dist is not None:
                cx.dist.barrier()
            torch.cuda.synchronize()
            l0 = sum(st.launch_count() for st in sts[:nl])
            ev0 = torch.cuda.Event(enable_timing=True)
            ev0.record(ext[0])
            for f in range(nf):
                for st in sts[:nl]:
                    st.render(0, 0, frame_index=(f + 1) % 256)
            ends = []
            for e in ext[:nl]:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(e)
                ends.append(ev)
            for st in sts[:nl]:
                st.finish()
            torch.cuda.synchronize()
            ms = max(ev0.elapsed_time(ev) for ev in ends)
            if cx.dist is not None:
                t = torch.tensor([ms], device="cuda")
                cx.dist.all_reduce(t, op=cx.dist.ReduceOp.MAX)
                ms = float(t.item())
            return ms, sum(st.launch_count() for st in sts[:nl]) - l0

        ms1, _ = run(1, 2)
        msn, launches = run(lanes, frames_per_lane)
    finally:
        for st in sts:
            st.close()
    steps_per_frame = wl["dst"][0] + (wl["dst"][1] - 1) * 2  # sierra-lite: shift 2
    fps1 = cx.world * 2 / (ms1 / 1e3)
    fpsn = cx.world * lanes * frames_per_lane / (msn / 1e3)
    return {"value": round(fpsn, 2), "unit": "frames/s", "frames_in_flight_per_gpu": lanes,
            "frames_per_lane": frames_per_lane, "ms_total": round(msn, 2),
            "one_frame_at_a_time": {"value": round(fps1, 2), "unit": "frames/s",
                                    "ms_per_frame": round(ms1 / 2, 2)},
            "error_diffusion_steps_per_frame": steps_per_frame,
            "column_steps_per_s": round(fpsn * steps_per_frame, 0),
            "gpu_launches": launches,
            "config": dict(workload_config(wl_name, cx.world), lanes=lanes),
            "note": "latency-bound: the error-diffusion pass is one 8-CTA cluster (8 SMs) per frame; frames in "
                    "flight are the unit of parallelism (SURVEY.md 8e: not shardable within a frame)"}


def roofline_of(res, wl_name, hbm_peak):
    plist = []
    for cnt, ns, desc in res["passes"]:
        b, label = pass_bytes(desc, res["wl"])
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
        traffic = tdb.get(wl_name, {}).get(top["pass"] if top else "", {}).get("dram_bytes")
    except Exception:  # noqa: BLE001
        pass
    if top and top["GBps"]:
        roof = {"bound": "hbm", "achieved": round(top["GBps"], 1), "peak": hbm_peak, "unit": "GB/s",
                "frac": round(top["GBps"] / hbm_peak, 4), "traffic": traffic,
                "traffic_source": "profiles/traffic.json (ncu --set full capture of the same kernel, "
                                  "committed; not measured in this run)",
                "kernel": top["pass"], "mean_us": round(top["mean_us"], 2),
                "algorithmic_bytes_per_launch": int(top["algorithmic_MB"] * 1e6),
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)",
                "note": "instruction-issue bound (fp32 + integer), not HBM bound; see DESIGN.md"}
    # whole frame by SURVEY.md 8(d): compulsory reads + writes of the frame / frame time
    return roof, plist


def parity_block(wl_name, frame):
    import numpy as np
    P, img, tgt, kw, grain = oracle_render_args(wl_name)
    k = dict(kw, frame_index=0)
    if grain:
        k["film_grain"] = (1, 1)
    want, _ = P.render(img, tgt, **k)
    if frame.dtype == np.uint16 and tgt.fmt == "rgba16hf":
        import plbtest as T
        d = T.ulp_diff_half(frame.view(np.float16), want)
        unit = "ULP of the fp16 target"
    else:
        d = np.abs(frame[..., :3].astype(np.int64) - want[..., :3].astype(np.int64))
        unit = "LSB of the target"
    return {"frame": f"{wl_name}, synthetic frame 0, fresh renderer", "max_diff": int(d.max()),
            "differing_fraction": float(np.mean(d > 0)), "unit": unit,
            "against": "CPU oracle (tests/pipeline_oracle.py over oracle/libpl_oracle.so)"}


def run_ours(args, rank, world, local_rank):
    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libplacebo_b200 has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    prev_affinity = numa_pin(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    cx = Ctx()
    cx.torch, cx.dist, cx.rank, cx.world, cx.local_rank = torch, dist, rank, world, local_rank
    cx.want_parity = not args.no_extra

    steps, warmup = args.steps, max(args.warmup, 3)
    res = measure(cx, args.workload, steps, warmup)
    extras = {}
    if not args.no_extra:
        for name in EXTRA:
            if name != args.workload:
                extras[name] = measure(cx, name, min(steps, 100), warmup)
    c5 = None
    if not args.no_extra and args.c5_lanes > 0:
        try:
            c5 = measure_lanes(cx, "c5_8k_dovi_errdiff", args.c5_lanes, 2)
        except Exception as e:  # noqa: BLE001
            c5 = {"error": str(e)}
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if prev_affinity is not None:
        os.sched_setaffinity(0, prev_affinity)  # the CPU arm below uses every core again
    if rank != 0:
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6560.6))

    def summarise(name, r):
        fps = world * r["steps"] / (r["ms"] / 1e3)
        roof, plist = roofline_of(r, name, hbm_peak)
        frame_bytes = r["in_bytes"] * (2 if r["wl"]["opts"].get("peak") else 1) + r["out_bytes"]
        d = {"value": round(fps, 2), "unit": "frames/s", "steps": r["steps"],
             "ms_per_step": round(r["ms"] / r["steps"], 4), "roofline": roof, "passes": plist,
             "frame_roofline": {"algorithmic_bytes_per_frame": frame_bytes,
                                "achieved_GBps": round(frame_bytes * fps / world / 1e9, 1),
                                "frac": round(frame_bytes * fps / world / 1e9 / hbm_peak, 4),
                                "note": "SURVEY.md 8(d): compulsory reads + writes of one frame / frame time"},
             "gpu_launches": r["launches"], "clocks": r["clocks"],
             "l2": f"ring of {r['ring']} resident input/output texture sets "
                   f"({r['ring'] * (r['in_bytes'] + r['out_bytes']) / 1e6:.0f} MB > 126 MB L2)"}
        if "ms_e2e" in r:
            d["e2e"] = {"value": round(world * r["e2e_steps"] / (r["ms_e2e"] / 1e3), 2), "unit": "frames/s",
                        "h2d_bytes_per_step": r["in_bytes"], "d2h_bytes_per_step": r["out_bytes"],
                        "steps": r["e2e_steps"], "ms_per_step": round(r["ms_e2e"] / r["e2e_steps"], 4)}
        if r.get("parity_frame") is not None:
            try:
                d["parity"] = parity_block(name, r["parity_frame"])
            except Exception as e:  # noqa: BLE001
                d["parity"] = {"error": str(e)}
        return d

    main = summarise(args.workload, res)
    cfg = workload_config(args.workload, world)
    cfg["l2"] = main.pop("l2")
    line = {
        "metric": METRIC, "value": main["value"], "unit": "frames/s",
        "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": main["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "roofline": main["roofline"], "frame_roofline": main["frame_roofline"],
        "passes": main["passes"],
        "e2e": main.get("e2e"),
        "gpu_launches": main["gpu_launches"],
        "clocks": main["clocks"],
    }
    if "parity" in main:
        line["parity"] = main["parity"]
    if extras:
        line["extra_workloads"] = {}
        for name, r in extras.items():
            d = summarise(name, r)
            d["config"] = dict(workload_config(name, world), l2=d.pop("l2"))
            line["extra_workloads"][name] = d
    if c5 is not None:
        line.setdefault("extra_workloads", {})["c5_8k_dovi_errdiff"] = c5
    if world == 1 and not args.no_cpu_baseline:
        try:
            t1, cores = cpu_frames(args.workload, 1, warm=1)
            n = max(5, min(15, int(20.0 / max(t1[0], 1e-3))))  # about 20 s of CPU work, >= 5 frames
            times, cores = cpu_frames(args.workload, n, warm=0)
            line["cpu_baseline"] = cpu_arm_line(args.workload, times, cores)
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
    if rank == 0:
        pin_openmp()  # before libgomp is loaded (oracle library)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        emit({"error": f"--gpus {args.gpus} needs torchrun (one rank per GPU)"})
        sys.exit(2)
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
