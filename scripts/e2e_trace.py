"""Ad-hoc: host-side timeline of one steady-state e2e generate call (DNAF_TRACE=1)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from dna_factory_b200 import _native
R = bench.ROWS_PER_STEP
NS = 14
sex, ctl, table, orow, osamp = bench.synth_population(NS * R, 0, window=R)
arrays = table.device_arrays()
def batch(k):
    lo, hi = k * R, (k + 1) * R
    p0, p1 = int(arrays["prefix_off"][lo]), int(arrays["prefix_off"][hi])
    return dict(chrom_class=arrays["chrom_class"][lo:hi], n_alleles=arrays["n_alleles"][lo:hi],
                thresholds=arrays["thresholds"][lo:hi], prefix_bytes=arrays["prefix_bytes"][p0:p1 + 1],
                prefix_off=arrays["prefix_off"][lo:hi + 1] - np.uint64(p0))
eng = _native.Engine(0)
out = torch.empty(400 << 20, dtype=torch.uint8, pin_memory=True).numpy()
eng.set_samples(sex, ctl); eng.set_chunk_bytes(bench.E2E_CHUNK)
zo, zs = np.zeros(0, np.uint64), np.zeros(0, np.uint32)
for k in range(NS):
    if k in (3, NS - 1): sys.stderr.write("==== traced step %d\n" % k); sys.stderr.flush()
    if k == 4: sys.stderr.write("==== end\n"); sys.stderr.flush()
    eng.set_snps(**batch(k)); eng.set_overrides(zo, zs)
    t0 = time.perf_counter(); eng.generate_into(0, R, 1, out, level=2); t1 = time.perf_counter()
print("last gen %.2f ms" % (1e3 * (t1 - t0)))
