"""The Python stream front-end (libplacebo_b200/stream.py over the ctypes glue) and the
asynchronous transfer engine: a software-pipelined loop (upload of frame i+1 and download of
frame i-1 overlapping the render of frame i on separate copy-engine streams) must produce
exactly the frames a fully serialised loop produces."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def small_workload():
    from libplacebo_b200 import stream as S
    wl = dict(S.WORKLOADS["c3_4k_hdr10_to_sdr"])
    wl["src"], wl["dst"] = (512, 288), (512, 288)
    return wl


def test_pipelined_stream_equals_serial():
    import torch
    from libplacebo_b200 import stream as S
    wl = small_workload()
    nframes, ring = 7, 3
    frames = [S.synth_yuv420p10(512, 288, seed=s) for s in range(nframes)]
    pinned = [[torch.from_numpy(p).pin_memory() for p in f] for f in frames]

    # serial: every step waits for the previous one
    st = S.Stream(device=0).configure(wl, ring=1)
    serial = []
    for i in range(nframes):
        st.upload(0, *pinned[i])
        st.render(0, 0, frame_index=i)
        out = torch.empty((288, 512, 4), dtype=torch.uint8).pin_memory()
        st.download(0, out)
        st.finish()
        serial.append(out.numpy().copy())
    peak_serial = st.get_peak()
    st.close()

    # pipelined: no host synchronisation inside the loop, ring of 3 slots reused round-robin
    st = S.Stream(device=0).configure(wl, ring=ring)
    outs = [torch.empty((288, 512, 4), dtype=torch.uint8).pin_memory() for _ in range(nframes)]
    for i in range(nframes):
        st.upload(i % ring, *pinned[i])
        st.render(i % ring, i % ring, frame_index=i)
        st.download(i % ring, outs[i])
    st.finish()
    assert st.get_peak() == peak_serial
    for i in range(nframes):
        assert np.array_equal(outs[i].numpy(), serial[i]), i
    assert st.launch_count() > 0
    st.close()


def test_stream_matches_pipeline_oracle():
    """End to end through the Python front-end against the CPU restatement (C3 at 512x288)."""
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    import pipeline_oracle as P
    import plbtest as T
    from libplacebo_b200 import stream as S
    wl = small_workload()
    planes = S.synth_yuv420p10(512, 288, seed=3)
    st = S.Stream(device=0).configure(wl, ring=1)
    st.upload(0, *planes)
    st.render(0, 0, frame_index=0)
    out = st.out_array()
    st.download(0, out)
    st.finish()
    st.close()
    img = P.Frame(planes, S.SYS[wl["src_sys"]], csp=np.array(wl["src_csp"], np.float32))
    tgt = P.Target(512, 288, "rgba8", sample_depth=8, color_depth=8, csp=np.array(wl["dst_csp"], np.float32))
    want, _ = P.render(img, tgt, upscaler="ewa_lanczos", deband=True, peak=2, dither=1, cmap_hq=True)
    assert np.array_equal(out[..., :3], want[..., :3])


def test_frame_parallel_flow_equals_sequential_stream():
    """The frame-parallel scheme on ONE GPU with two simulated ranks (measure -> exchange ->
    replay -> render, libplacebo_b200/parallel.py) must reproduce the frames of a plain
    sequential pl_render_image stream bit for bit, including the smoothed peak state."""
    import torch
    from libplacebo_b200 import parallel
    from libplacebo_b200 import stream as S
    wl = small_workload()
    nframes, world = 6, 2
    frames = [S.synth_yuv420p10(512, 288, seed=40 + s) for s in range(nframes)]
    frames[3] = tuple(np.clip(p.astype(np.int32) // 3 + 64, 0, 1023).astype(np.uint16) for p in frames[3])  # scene change

    st = S.Stream(device=0).configure(wl, ring=1)
    seq, seq_state = [], []
    for i in range(nframes):
        st.upload(0, *frames[i])
        st.render(0, 0, frame_index=i)
        out = st.out_array()
        st.download(0, out)
        st.finish()
        seq.append(out.copy())
        seq_state.append(st.get_peak())
    st.close()

    # two "ranks" sharing this process: rank r handles frames r, r + 2, ...
    ranks = [S.Stream(device=0).configure(wl, ring=1) for _ in range(world)]
    state = (0.0, 0.0)
    for step in range(nframes // world):
        bufs = []
        for r in range(world):
            f = step * world + r
            ranks[r].upload(0, *frames[f])
            bufs.append(ranks[r].measure_peak(0, f))
        for r in range(world):
            mine, last = parallel.replay(ranks[r].peak_update, state, bufs, r)
            ranks[r].set_peak(mine)
            f = step * world + r
            ranks[r].render(0, 0, frame_index=f)
            out = ranks[r].out_array()
            ranks[r].download(0, out)
            ranks[r].finish()
            assert np.array_equal(out, seq[f]), f
            assert np.allclose(mine, seq_state[f], rtol=0, atol=0), (f, mine, seq_state[f])
        state = last
    for r in ranks:
        r.close()
