"""Frame-parallel rendering of ONE stream across GPUs (one process per GPU, torch.distributed).

The path's only cross-frame state is the peak-detection IIR (shaders/colorspace.c:1071-1180 in
the reference): frame f's tone curve depends on the smoothed peak after frames 0..f. Frames are
dealt round-robin (frame f = step * world + rank). Per step every rank

  1. measures its frame: the pass graph up to the peak-detection tap (`pl_renderer_measure_peak`:
     raw `peak_buf_data`, 3264 bytes; the rgba16hf image of that pass stays cached in the renderer,
     it is exactly what the sequential same-frame flow stores between its two passes);
  2. all-gathers the statistics (NCCL over NVLink: 3264 B per rank — the one collective of the
     path; there is no pixel exchange);
  3. replays the recurrence over the gathered buffers in frame order
     (`pl_peak_state_update`, host arithmetic, identical on every rank), stopping at its own
     frame to obtain the state the sequential reference would have had;
  4. renders its frame with that state (`pl_renderer_set_peak_state` + `pl_render_image`, which
     continues from the cached image: the source pass is not repeated).

The output is bit-identical to rendering the stream on one GPU (tests/test_parallel.py checks
the replay logic with the gloo backend on CPU, tests/test_gpu_parallel.py on GPUs).
"""
import numpy as np

from .stream import PEAK_BUF_SIZE


def replay(states_fn, state, bufs, upto):
    """Apply the recurrence to bufs[0..upto] starting from `state`; returns the state after
    bufs[upto] and the state after all of them. `states_fn(state, buf)` -> (new_state, used)."""
    mine = None
    for i, b in enumerate(bufs):
        state, _ = states_fn(state, b)
        if i == upto:
            mine = state
    return mine, state


class PeakExchange:
    def __init__(self, stream, dist, rank, world, device=None):
        import torch
        self.st, self.dist, self.rank, self.world = stream, dist, rank, world
        self.torch = torch
        self.state = (0.0, 0.0)  # smoothed (avg_pq, max_pq) after the last completed step
        backend = dist.get_backend()
        self.dev = device if device is not None else ("cuda" if backend == "nccl" else "cpu")
        self.send = torch.zeros(PEAK_BUF_SIZE // 4, dtype=torch.int32, device=self.dev)
        self.recv = torch.zeros(world * PEAK_BUF_SIZE // 4, dtype=torch.int32, device=self.dev)

    def exchange(self, buf):
        """buf: uint32[816] statistics of this rank's frame -> list of all ranks' buffers."""
        t = self.torch
        self.send.copy_(t.from_numpy(buf.view(np.int32)))
        self.dist.all_gather_into_tensor(self.recv, self.send)
        allb = self.recv.cpu().numpy().view(np.uint32).reshape(self.world, -1)
        return [allb[r] for r in range(self.world)]

    def render(self, slot, out_slot, step):
        frame_index = (step * self.world + self.rank) % 256
        buf = self.st.measure_peak(slot, frame_index)
        bufs = self.exchange(buf)
        mine, last = replay(self.st.peak_update, self.state, bufs, self.rank)
        self.state = last
        self.st.set_peak(mine)
        self.st.render(slot, out_slot, frame_index=frame_index)


class PeakPipeline:
    """The same scheme with the host latency hidden (CUDA ranks, NCCL): while the statistics of
    frame i are all-gathered on a side stream and replayed on the host, the GPU already measures
    frame i+1 on the library's stream; only the few-hundred-byte exchange is ever waited for,
    never the render queue.

        pipe = PeakPipeline(stream, dist, rank, world, ring)
        for i in range(n): pipe.render(i)          # frame i of this rank, slot = i % ring
        pipe.drain()

    `before_measure(j)` (optional) runs right before frame j is measured — bench.py uploads the
    frame's planes there in its end-to-end mode."""

    def __init__(self, stream, dist, rank, world, ring, before_measure=None):
        import torch
        self.st, self.dist, self.rank, self.world, self.ring = stream, dist, rank, world, ring
        self.torch = torch
        self.before = before_measure
        self.state = (0.0, 0.0)
        self.main = torch.cuda.ExternalStream(stream.cuda_stream)
        self.side = torch.cuda.Stream()
        n = PEAK_BUF_SIZE // 4
        self.send = [torch.zeros(n, dtype=torch.int32, device="cuda") for _ in range(2)]
        self.recv = [torch.zeros(world * n, dtype=torch.int32, device="cuda") for _ in range(2)]
        self.host = [torch.zeros(world * n, dtype=torch.int32).pin_memory() for _ in range(2)]
        self.ev_meas = [torch.cuda.Event() for _ in range(2)]
        self.ev_xchg = [torch.cuda.Event() for _ in range(2)]
        self.inflight = None   # index of the frame whose measurement / exchange is in flight

    def _frame_index(self, i):
        return (i * self.world + self.rank) % 256

    def _submit(self, i):
        b = i & 1
        if self.before is not None:
            self.before(i)
        self.st.measure_peak_dev(i % self.ring, self._frame_index(i), self.send[b].data_ptr())
        self.ev_meas[b].record(self.main)
        with self.torch.cuda.stream(self.side):
            self.side.wait_event(self.ev_meas[b])
            self.dist.all_gather_into_tensor(self.recv[b], self.send[b])
            # kernel copy into the pinned buffer: a DMA copy of these few KB would queue behind the
            # frame-sized downloads on the device-to-host copy engine
            self.st.small_copy(self.host[b].data_ptr(), self.recv[b].data_ptr(),
                               self.recv[b].numel() * 4, self.side.cuda_stream)
            self.ev_xchg[b].record(self.side)
        self.inflight = i

    def render(self, i, out_slot=None, last=False):
        if self.inflight != i:
            self._submit(i)
        b = i & 1
        if not last:
            self._submit(i + 1)          # the GPU has work while the host waits below
            self.inflight = i + 1
        else:
            self.inflight = None
        self.ev_xchg[b].synchronize()     # waits for the exchange of frame i only
        allb = self.host[b].numpy().view(np.uint32).reshape(self.world, -1)
        mine, last_state = replay(self.st.peak_update, self.state, [allb[r] for r in range(self.world)],
                                  self.rank)
        self.state = last_state
        self.st.set_peak(mine)
        slot = i % self.ring
        self.st.render(slot, slot if out_slot is None else out_slot, frame_index=self._frame_index(i))

    def drain(self):
        """Forget a measurement submitted ahead (its image simply expires in the renderer)."""
        self.side.synchronize()
        self.inflight = None
