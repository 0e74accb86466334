"""Ad-hoc: wall time per generate_device call vs CUDA-event time (host overhead of a pass)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from dna_factory_b200 import _native
R = bench.ROWS_PER_STEP
sex, ctl, table, orow, osamp = bench.synth_population(8 * R, 0, window=R)
eng = _native.Engine(0)
eng.set_samples(sex, ctl); eng.set_snps(**table.device_arrays()); eng.set_overrides(orow, osamp)
for k in range(8):
    t0 = time.perf_counter(); st = eng.generate_device(k * R, (k + 1) * R, 1, level=2); t1 = time.perf_counter()
    print("step %d wall %.3f ms, events total %.3f ms (fused %.3f, deflate %.3f)" % (k, 1e3 * (t1 - t0), st["ms_total"], st["ms_fused"], st["ms_deflate"]))
