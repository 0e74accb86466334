"""Ad-hoc timing of the end-to-end step pieces (host set-up vs generate) for several pass sizes."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from dna_factory_b200 import _native

R = bench.ROWS_PER_STEP
sex, ctl, table, orow, osamp = bench.synth_population(6 * R, 0, window=R)
arrays = table.device_arrays()
eng = _native.Engine(0)
eng.set_samples(sex, ctl)
out = torch.empty(400 << 20, dtype=torch.uint8, pin_memory=True).numpy()

def batch(k):
    lo, hi = k * R, (k + 1) * R
    p0, p1 = int(arrays["prefix_off"][lo]), int(arrays["prefix_off"][hi])
    return dict(chrom_class=arrays["chrom_class"][lo:hi], n_alleles=arrays["n_alleles"][lo:hi],
                thresholds=arrays["thresholds"][lo:hi], prefix_bytes=arrays["prefix_bytes"][p0:p1 + 1],
                prefix_off=arrays["prefix_off"][lo:hi + 1] - np.uint64(p0))

for chunk in (1 << 30, 320 << 20, 160 << 20, 80 << 20, 40 << 20):
    eng.set_chunk_bytes(chunk)
    ts, tg, tp = [], [], []
    for k in range(6):
        b = batch(k)
        t0 = time.perf_counter(); eng.set_snps(**b); eng.set_overrides(np.zeros(0, np.uint64), np.zeros(0, np.uint32))
        t1 = time.perf_counter(); eng.plan(0, R)
        t2 = time.perf_counter(); st = eng.generate_into(0, R, 1, out, level=2)
        t3 = time.perf_counter()
        ts.append(t1 - t0); tp.append(t2 - t1); tg.append(t3 - t2)
    print("chunk %4d MB: set_snps %.2f ms, layout/tables %.2f ms, generate %.2f ms (kernels %.2f ms, %d MB out)" % (
        chunk >> 20, 1e3 * np.median(ts[2:]), 1e3 * np.median(tp[2:]), 1e3 * np.median(tg[2:]), st["ms_total"], st["bgzf_bytes"] >> 20))
t0 = time.perf_counter(); dev = eng.generate_device(0, R, 1, level=2); print("device-only %.2f ms" % (1e3 * (time.perf_counter() - t0)))
