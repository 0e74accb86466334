"""Pinned D2H bandwidth with ONE PROCESS PER GPU (launch under torchrun): what bounds the end-to-end rate at N > 1.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/d2h_probe_mp.py

For every variant -- cudaHostAlloc default / portable / write-combined, 1 or 2 copy streams -- all ranks copy at once
(barrier before, wall clock max over ranks) and rank 0 prints per-GPU and aggregate GB/s.  Also prints the solo rate of
rank 0 (others idle).  cuda-python's runtime bindings are used so that the allocation flags are under control
(torch's pin_memory is cudaHostAlloc(default))."""
import os, time
import torch, torch.distributed as dist
from cuda import cudart

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
N = 256 << 20
dev = torch.empty(N, dtype=torch.uint8, device="cuda"); dev.fill_(7)

def ck(r):
    if isinstance(r, tuple):
        if int(r[0]) != 0: raise RuntimeError("cuda error %s" % r[0])
        return r[1] if len(r) == 2 else r[1:]
    if int(r) != 0: raise RuntimeError("cuda error %s" % r)

def barrier():
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()

def run(flags, n_streams, active=True, reps=8):
    host = ck(cudart.cudaHostAlloc(N, flags))
    streams = [ck(cudart.cudaStreamCreateWithFlags(cudart.cudaStreamNonBlocking)) for _ in range(n_streams)]
    piece = N // n_streams
    def go():
        for _ in range(reps):
            for i, s in enumerate(streams):
                ck(cudart.cudaMemcpyAsync(host + i * piece, dev.data_ptr() + i * piece, piece, cudart.cudaMemcpyKind.cudaMemcpyDeviceToHost, s))
        for s in streams: ck(cudart.cudaStreamSynchronize(s))
    if active: go()
    barrier()
    t0 = time.perf_counter()
    if active: go()
    dt = time.perf_counter() - t0
    barrier()
    for s in streams: ck(cudart.cudaStreamDestroy(s))
    ck(cudart.cudaFreeHost(host))
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return reps * N / dt / 1e9, float(t.item())

variants = [("default", cudart.cudaHostAllocDefault), ("portable", cudart.cudaHostAllocPortable), ("write-combined", cudart.cudaHostAllocWriteCombined)]
for name, fl in variants:
    for ns in (1, 2):
        mine, tmax = run(fl, ns)
        agg = world * 8 * N / tmax / 1e9
        g = torch.tensor([mine], dtype=torch.float64, device="cuda"); allg = [torch.zeros_like(g) for _ in range(world)]
        if world > 1: dist.all_gather(allg, g)
        else: allg = [g]
        if rank == 0:
            print("%-15s %d stream(s): aggregate %.1f GB/s over %d GPUs; per GPU %s" % (name, ns, agg, world, " ".join("%.1f" % float(x.item()) for x in allg)), flush=True)
mine, _ = run(cudart.cudaHostAllocDefault, 1, active=(rank == 0))
if rank == 0:
    print("rank 0 alone (others idle): %.1f GB/s" % mine, flush=True)
if world > 1: dist.destroy_process_group()
