"""Ad-hoc: e2e steps with 1 vs 2 contexts, per-step wall times."""
import sys, time, os, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from dna_factory_b200 import _native

R = bench.ROWS_PER_STEP
NS = 36
sex, ctl, table, orow, osamp = bench.synth_population(NS * R, 0, window=R)
arrays = table.device_arrays()

def batch(k):
    lo, hi = k * R, (k + 1) * R
    p0, p1 = int(arrays["prefix_off"][lo]), int(arrays["prefix_off"][hi])
    return dict(chrom_class=arrays["chrom_class"][lo:hi], n_alleles=arrays["n_alleles"][lo:hi],
                thresholds=arrays["thresholds"][lo:hi], prefix_bytes=arrays["prefix_bytes"][p0:p1 + 1],
                prefix_off=arrays["prefix_off"][lo:hi + 1] - np.uint64(p0))
batches = [batch(k) for k in range(NS)]
zo, zs = np.zeros(0, np.uint64), np.zeros(0, np.uint32)

for n_ctx in (1, 2, 3):
    engines = [_native.Engine(0) for _ in range(n_ctx)]
    outs = [torch.empty(400 << 20, dtype=torch.uint8, pin_memory=True).numpy() for _ in engines]
    for e in engines:
        e.set_samples(sex, ctl); e.set_chunk_bytes(bench.E2E_CHUNK)
    log = []
    def worker(j, ks):
        for k in ks[j::n_ctx]:
            t0 = time.perf_counter(); engines[j].set_snps(**batches[k]); engines[j].set_overrides(zo, zs)
            t1 = time.perf_counter(); st = engines[j].generate_into(0, R, 1, outs[j], level=2)
            t2 = time.perf_counter(); log.append((k, j, t0, t1, t2))
    def run(ks):
        ts = [threading.Thread(target=worker, args=(j, ks)) for j in range(n_ctx)]
        [t.start() for t in ts]; [t.join() for t in ts]
    run(list(range(4)))
    log.clear()
    torch.cuda.synchronize(); T0 = time.perf_counter()
    run(list(range(4, NS)))
    torch.cuda.synchronize(); T1 = time.perf_counter()
    print("n_ctx %d: %.2f ms per step" % (n_ctx, 1e3 * (T1 - T0) / (NS - 4)))
    L = sorted(log)
    for i in range(0, len(L), 8):
        seg = L[i:i + 8]
        print("   steps %2d-%2d: %.2f ms per step (set %.2f, gen %.2f)" % (seg[0][0], seg[-1][0], 1e3 * (max(x[4] for x in seg) - min(x[2] for x in seg)) / len(seg),
              1e3 * np.mean([x[3] - x[2] for x in seg]), 1e3 * np.mean([x[4] - x[3] for x in seg])))
    del engines
