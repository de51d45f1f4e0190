"""Raw PCIe rates of the box (pinned host memory): one direction alone and both at once --
the ceiling of bench.py's e2e arm (HostStepper moves 60 B per particle each way per timestep)."""
import torch, time
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both():
    h2d(); d2h()
print("H2D alone  %.1f GB/s" % (n / timed(h2d) / 1e9))
print("D2H alone  %.1f GB/s" % (n / timed(d2h) / 1e9))
print("both       %.1f GB/s each way" % (n / timed(both) / 1e9))
