"""Ad-hoc: kernel-only throughput (generate_device) for another population shape: python scripts/shape_probe.py N ROWS."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from dna_factory_b200 import _native
N = int(sys.argv[1]); R = int(sys.argv[2])
bench.N_CASES = N // 2; bench.N_CONTROLS = N - N // 2
sex, ctl, table, orow, osamp = bench.synth_population(6 * R, 0, window=R)
eng = _native.Engine(0)
eng.set_samples(sex, ctl); eng.set_snps(**table.device_arrays()); eng.set_overrides(orow, osamp)
for k in range(6):
    t0 = time.perf_counter(); st = eng.generate_device(k * R, (k + 1) * R, 1, level=2); t1 = time.perf_counter()
    if k >= 2:
        print("N %d rows %d: wall %.3f ms -> %.3e calls/s; fused %.3f ms (%.0f GB/s of text = %.3f of 6546.9), ratio %.2f" % (
            N, R, 1e3 * (t1 - t0), st["calls"] / (t1 - t0), st["ms_fused"], st["text_bytes"] / st["ms_fused"] / 1e6,
            st["text_bytes"] / st["ms_fused"] / 1e6 / 6546.9, st["text_bytes"] / st["bgzf_bytes"]))
