"""Ad-hoc: aggregate pinned D2H bandwidth with every visible GPU copying at once (what bounds e2e at N > 1)."""
import time, threading, torch
n_gpu = torch.cuda.device_count()
n = 256 << 20
bufs = []
for g in range(n_gpu):
    with torch.cuda.device(g):
        bufs.append((torch.empty(n, dtype=torch.uint8, device="cuda:%d" % g), torch.empty(n, dtype=torch.uint8, pin_memory=True),
                     torch.cuda.Stream(device=g)))
def run(active):
    for g in active:
        d, h, s = bufs[g]
        with torch.cuda.stream(s):
            for _ in range(8):
                h.copy_(d, non_blocking=True)
    for g in active:
        bufs[g][2].synchronize()
for k in range(1, n_gpu + 1):
    active = list(range(k))
    run(active)
    t0 = time.perf_counter(); run(active); dt = time.perf_counter() - t0
    print("%d GPUs copying at once: %.1f GB/s aggregate D2H" % (k, k * 8 * n / dt / 1e9))
