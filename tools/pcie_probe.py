"""PCIe probe (GPU box): pinned H2D / D2H bandwidth alone and concurrently, 1-D and pitched 2-D."""
import torch, time
dev = torch.device("cuda:0")
n = 64 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device=dev)
d_b = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(f, reps=20):
    f(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps
def h2d():
    with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
def both():
    h2d(); d2h()
for name, f, nb in (("H2D", h2d, n), ("D2H", d2h, n), ("H2D+D2H concurrently", both, 2 * n)):
    dt = run(f)
    print(f"{name}: {nb / dt / 1e9:.1f} GB/s total ({dt * 1e3:.2f} ms per {n >> 20} MiB each)")
import subprocess
print(subprocess.run(["nvidia-smi", "--query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max", "--format=csv"], capture_output=True, text=True).stdout)
print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[:1500])
