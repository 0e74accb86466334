"""world_size-2 gloo test of the multi-rank plumbing (CPU): contiguous SNP ranges, max/sum reductions, and
rank streams that concatenate into the single-rank result (checked with the CPU oracle)."""
import os

import numpy as np
import torch.multiprocessing as mp

from tests.cases import synth_case


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from dna_factory_b200 import partition
    from oracle import oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    case = synth_case(120, 31, seed=5)
    flat = oracle.flatten(case.samples, case.snps)
    row_bytes = np.diff(oracle.rows_from_flat(flat, case.seed)[1].astype(np.int64))
    bounds = partition.row_bounds(len(case.snps), world, row_bytes)
    lo, hi = bounds[rank], bounds[rank + 1]
    text, _ = oracle.rows(case.samples, case.snps[lo:hi], case.seed, lo)   # rows are global: row_begin = lo
    blob = oracle.bgzf(text, level=2, with_eof=False)
    # the cross-rank parity check bench.py runs under NCCL: every rank publishes the signature of the first rows of its
    # own range; rank 0 regenerates each of them alone (here: with the oracle) and compares
    W = 3
    head, _ = oracle.rows(case.samples, case.snps[lo:lo + W], case.seed, lo)
    hb = oracle.bgzf(head, level=2, with_eof=False)
    _, n_blocks, _ = oracle.bgzf_decompress(hb)

    def recompute(r):
        t, _ = oracle.rows(case.samples, case.snps[bounds[r]:bounds[r] + W], case.seed, bounds[r])
        return partition.window_signature(t, oracle.bgzf_decompress(oracle.bgzf(t, level=2, with_eof=False))[1], 7)

    verdict = partition.check_rank_windows(partition.window_signature(head, n_blocks, 7), recompute, dist)
    # a rank that drew its rows with the wrong row base must be caught
    wrong, _ = oracle.rows(case.samples, case.snps[lo:lo + W], case.seed, lo + (1 if rank == 1 else 0))
    verdict_bad = partition.check_rank_windows(partition.window_signature(wrong, n_blocks, 7), recompute, dist)
    slowest = partition.reduce_max(float(rank + 1), dist)
    total = partition.reduce_sum(float(hi - lo), dist)
    dist.barrier()
    q.put((rank, bounds, blob, slowest, total, verdict, verdict_bad))
    dist.destroy_process_group()


def test_two_ranks_concatenate_to_single_rank_stream():
    from oracle import oracle
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, b0, blob0, mx0, tot0, v0, vb0), (r1, b1, blob1, mx1, tot1, v1, vb1) = got
    assert v0 == "ok" and v1 is None and vb0.startswith("FAILED: rank 1") and vb1 is None
    assert b0 == b1 and b0[0] == 0 and b0[-1] == 31 and 0 < b0[1] < 31
    assert mx0 == mx1 == 2.0 and tot0 == tot1 == 31.0
    case = synth_case(120, 31, seed=5)
    want, _ = oracle.rows(case.samples, case.snps, case.seed, 0)
    assert oracle.bgzf_decompress(blob0 + blob1)[0] == want


def test_row_bounds_balance_by_bytes():
    from dna_factory_b200 import partition
    assert partition.row_bounds(10, 4) == [0, 2, 5, 7, 10]
    b = partition.row_bounds(6, 2, [100, 100, 100, 100, 10, 10])
    assert b == [0, 3, 6] or b == [0, 2, 6]
    assert partition.row_bounds(3, 8)[-1] == 3
