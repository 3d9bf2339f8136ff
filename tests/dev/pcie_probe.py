"""Development aid: pinned-memory PCIe floor of this box (H2D, D2H, both at once), 1 GiB each."""
import torch, time
n = 1 << 30
h_a = torch.empty(n, dtype=torch.uint8).pin_memory(); h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda"); d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timed(f, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); f(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best
def h2d():
    with torch.cuda.stream(s1): d_a.copy_(h_a, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_b.copy_(d_b, non_blocking=True)
def both():
    h2d(); d2h()
for name, f in (("h2d", h2d), ("d2h", d2h), ("both", both)):
    t = timed(f); print(f"{name}: {1e3 * t:.2f} ms for 1 GiB each way -> {n / t / 1e9:.1f} GB/s per direction")
