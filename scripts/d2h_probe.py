"""Ad-hoc: pinned D2H / H2D bandwidth of the box for several transfer sizes (what bounds the e2e number)."""
import time, torch
for mb in (4, 16, 64, 256):
    n = mb << 20
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    for name, (dst, src) in (("D2H", (h, d)), ("H2D", (d, h))):
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(10):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 10
        print("%s %4d MB: %.1f GB/s" % (name, mb, n / dt / 1e9))
