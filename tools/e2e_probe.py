"""Where does the end-to-end frame time go? (GPU box) uploads / renders / downloads alone and combined."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from libplacebo_b200 import stream as S
wl = S.WORKLOADS["c3_4k_hdr10_to_sdr"]
sw, sh = wl["src"]; dw, dh = wl["dst"]
ring = 4
st = S.Stream(device=0).configure(wl, ring=ring)
lib = torch.cuda.ExternalStream(st.cuda_stream, device=0)
hin = []
for r in range(ring):
    y, u, v = S.synth_yuv420p10(sw, sh, seed=r)
    hin.append([torch.from_numpy(p).pin_memory() for p in (y, u, v)])
    st.upload(r, *hin[-1])
hout = [torch.empty((dh, dw, 4), dtype=torch.uint8).pin_memory() for _ in range(2)]
st.finish()
def timed(name, f, n=60):
    for i in range(5): f(i)
    st.join(); st.finish(); torch.cuda.synchronize()
    t = time.perf_counter()
    for i in range(n): f(i)
    st.join(); st.finish(); torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / n
    print(f"{name:34s} {dt*1e3:7.3f} ms/frame  {1/dt:8.1f} fps", flush=True)
timed("upload only", lambda i: st.upload(i % ring, *hin[i % ring]))
timed("render only", lambda i: st.render(i % ring, i % ring, frame_index=i % 256))
timed("download only", lambda i: st.download(i % ring, hout[i % 2]))
def ud(i):
    st.upload(i % ring, *hin[i % ring]); st.download((i + 1) % ring, hout[i % 2])
timed("upload + download (no dependency)", ud)
def full(i):
    s = i % ring
    st.upload(s, *hin[s]); st.render(s, s, frame_index=i % 256); st.download(s, hout[i % 2])
timed("upload + render + download", full)
def ur(i):
    s = i % ring
    st.upload(s, *hin[s]); st.render(s, s, frame_index=i % 256)
timed("upload + render", ur)
def rd(i):
    s = i % ring
    st.render(s, s, frame_index=i % 256); st.download(s, hout[i % 2])
timed("render + download", rd)
