"""Frame-parallel peak-state replay (libplacebo_b200/parallel.py) with torch.distributed, backend
gloo, world size 2, on the CPU. The GPU parts (measure / render) are replaced by a fake stream
that returns pre-computed statistics; what is under test is the host logic every rank runs:
all-gather of the 3264-byte buffers and the in-order replay of pl_peak_state_update (the
product's own host function, called through the test shim), which must reproduce the state
sequence of a single sequential stream exactly."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

from libplacebo_b200 import parallel  # noqa: E402
from libplacebo_b200.stream import PEAK_BUF_SIZE  # noqa: E402

SLICES, BINS = 12, 64


def synth_stats(seed, level):
    """A plausible struct peak_buf_data: 12 slices of wg_count, wg_active, sum, max, hist[64]."""
    r = np.random.default_rng(seed)
    buf = np.zeros(PEAK_BUF_SIZE // 4, np.uint32)
    wg = 8
    buf[0:SLICES] = wg
    buf[SLICES:2 * SLICES] = wg
    avg = int(level * 16383)
    buf[2 * SLICES:3 * SLICES] = wg * avg + r.integers(0, 50, SLICES)
    buf[3 * SLICES:4 * SLICES] = min(16383, int(avg * 1.3)) + r.integers(0, 20, SLICES)
    hist = buf[4 * SLICES:].reshape(SLICES, BINS)
    b = int(np.clip((avg >> 7) - 64 + 64, 0, 63))
    hist[:, max(b - 1, 0)] = 400
    hist[:, b] = 1200
    hist[:, min(b + 1, 63)] = 448
    return buf


class FakeStream:
    """Stands in for stream.Stream: statistics come from a table, peak_update is the product's
    pl_peak_state_update (host C code) reached through the test shim."""

    def __init__(self, frames):
        import plbtest as T
        self.frames = frames
        self.shim = T.shim()
        self.rendered = []
        self.state_for_render = None
        self.cursor = None

    def measure_peak(self, slot, frame_index=-1):
        return self.frames[slot].copy()

    def peak_update(self, state, buf):
        import ctypes as C
        st = np.array(state, np.float32)
        used = self.shim.shim_peak_update(st.ctypes.data_as(C.c_void_p), buf.ctypes.data_as(C.c_void_p), 1)
        return (float(st[0]), float(st[1])), bool(used)

    def set_peak(self, state):
        self.state_for_render = state

    def render(self, slot, out_slot, frame_index=-1):
        self.rendered.append((slot, self.state_for_render, frame_index))


def _worker(rank, world, port, nsteps, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    levels = [0.30, 0.31, 0.32, 0.70, 0.71, 0.30, 0.305, 0.31]
    frames = {f: synth_stats(f, levels[f % len(levels)]) for f in range(nsteps * world)}
    st = FakeStream(frames)
    ex = parallel.PeakExchange(st, dist, rank, world)
    for step in range(nsteps):
        ex.render(step * world + rank, 0, step)
    q.put((rank, st.rendered))
    dist.destroy_process_group()


def test_frame_parallel_replay_matches_sequential():
    world, nsteps = 2, 4
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nsteps, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    # sequential reference: one stream applying the recurrence frame by frame
    levels = [0.30, 0.31, 0.32, 0.70, 0.71, 0.30, 0.305, 0.31]
    seq = FakeStream({})
    state, want = (0.0, 0.0), []
    for f in range(nsteps * world):
        state, _ = seq.peak_update(state, synth_stats(f, levels[f % len(levels)]))
        want.append(state)
    for rank in range(world):
        assert len(got[rank]) == nsteps
        for step, (slot, st, fi) in enumerate(got[rank]):
            f = step * world + rank
            assert slot == f and fi == f % 256
            assert st == want[f], (rank, step, st, want[f])
    # the scene change (0.32 -> 0.70) must actually have exercised the non-trivial branch
    assert want[3][0] > want[2][0] * 1.5
