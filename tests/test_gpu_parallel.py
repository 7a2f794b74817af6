"""Frame-parallel rendering on real GPUs (NCCL): two ranks deal the frames of ONE stream
round-robin, exchange the peak statistics (libplacebo_b200/parallel.py, pipelined all-gather on a
side stream) and must reproduce, bit for bit, what a single sequential renderer produces — frames
AND the smoothed peak state (the only cross-frame state of the path, colorspace.c:1124-1152).

Needs two CUDA devices (`gpurun --gpus 2`); skipped on a single-GPU box.
"""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
W, H, FRAMES = 512, 288, 8


def _workload():
    from libplacebo_b200 import stream as S
    wl = dict(S.WORKLOADS["c3_4k_hdr10_to_sdr"])
    wl["src"] = wl["dst"] = (W, H)
    wl["dst_fmt"], wl["dst_depth"] = "rgba16", 16
    return wl


def _frame(f):
    from libplacebo_b200 import stream as S
    y, u, v = S.synth_yuv420p10(W, H, seed=f)
    # a scene change half way: exercises the non-trivial branch of the recurrence
    if f >= FRAMES // 2:
        y = np.clip(y.astype(np.int32) + 180, 0, 1023).astype(np.uint16)
    return y, u, v


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    import torch
    import torch.distributed as dist
    from libplacebo_b200 import parallel, stream as S
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    steps = FRAMES // world
    st = S.Stream(device=rank).configure(_workload(), ring=steps)
    for i in range(steps):
        st.upload(i, *_frame(i * world + rank))
    pipe = parallel.PeakPipeline(st, dist, rank, world, ring=steps)
    outs = []
    for i in range(steps):
        pipe.render(i, last=i == steps - 1)
        o = st.out_array()
        st.download(i, o)
        st.join()
        st.finish()
        outs.append((i * world + rank, o.copy(), pipe.state))
    pipe.drain()
    q.put((rank, outs))
    st.close()
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_stream_equals_sequential_stream():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import torch.multiprocessing as mp
    from libplacebo_b200 import stream as S

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        rank, outs = q.get(timeout=300)
        for f, o, state in outs:
            got[f] = (o, state)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0

    # the sequential reference: one renderer, frames in order, its own IIR state
    st = S.Stream(device=0).configure(_workload(), ring=FRAMES)
    for f in range(FRAMES):
        st.upload(f, *_frame(f))
    for f in range(FRAMES):
        st.render(f, f, frame_index=f % 256)
        want = st.out_array()
        st.download(f, want)
        st.join()
        st.finish()
        assert np.array_equal(got[f][0], want), f"frame {f} differs from the sequential stream"
    seq_state = st.get_peak()
    st.close()
    # after the last step every rank holds the state after ALL frames
    assert got[FRAMES - 1][1] == pytest.approx(seq_state, abs=0) or got[FRAMES - 1][1] == seq_state


def test_two_gpus_in_one_process_interleaved():
    """Two pl_gpu objects on two devices driven from ONE process, calls interleaved, while torch
    keeps a third notion of the current device: every entry point selects its own device
    (allocations, uploads, launches, TMA descriptors, per-device constants and curve tables).
    Both must produce the sequential single-GPU result."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    sys.path.insert(0, ROOT)
    from libplacebo_b200 import stream as S
    sts = [S.Stream(device=d).configure(_workload(), ring=2) for d in (0, 1)]
    outs = [[], []]
    try:
        for f in range(3):
            for d, st in enumerate(sts):
                torch.cuda.set_device(1 - d)          # the "wrong" device is current on purpose
                st.upload(f % 2, *_frame(f))
            for d, st in enumerate(sts):
                torch.cuda.set_device(d if f % 2 else 1 - d)
                st.render(f % 2, f % 2, frame_index=f)
                o = st.out_array()
                st.download(f % 2, o)
                st.join()
                st.finish()
                outs[d].append(o.copy())
    finally:
        for st in sts:
            st.close()
    for f in range(3):
        assert np.array_equal(outs[0][f], outs[1][f]), f"frame {f}: device 0 and device 1 differ"
    assert outs[0][0].any()
