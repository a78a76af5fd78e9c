"""Pinned host -> device copy bandwidth at the bench's step size (157 MB), alone and with a D2H in the other direction."""
import torch, time
n = 512 * 640 * 480
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(31459328, dtype=torch.uint8).pin_memory()
d2 = torch.empty(31459328, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for both in (False, True):
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        with torch.cuda.stream(s1):
            d.copy_(h, non_blocking=True)
        if both:
            with torch.cuda.stream(s2):
                h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 20
    print("h2d %s: %.3f ms per 157 MB = %.1f GB/s" % ("with d2h" if both else "alone", dt * 1e3, n / dt / 1e9))
