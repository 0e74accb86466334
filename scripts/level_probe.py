"""Ad-hoc: ratio and device-only rate of every -z level on the C2 shape (20 000 samples), one 32768-row window;
   python scripts/level_probe.py [rows] [samples]."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from dna_factory_b200 import _native
R = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
if len(sys.argv) > 2:                      # another sample count, e.g. 100000 for the C4 shape
    bench.N_CASES = int(sys.argv[2]) // 2
    bench.N_CONTROLS = int(sys.argv[2]) - bench.N_CASES
sex, ctl, table, orow, osamp = bench.synth_population(R, 0, window=R)
eng = _native.Engine(0)
eng.set_samples(sex, ctl); eng.set_snps(**table.device_arrays()); eng.set_overrides(orow, osamp)
for lv in (1, 2, 3, 4, 5, 6, 7, 8, 9):
    t0 = time.perf_counter()
    eng.generate_device(0, R, bench.PHILOX_SEED, level=lv)          # builds the tier's tables
    t_first = time.perf_counter() - t0
    best = None
    for _ in range(3):
        st = eng.generate_device(0, R, bench.PHILOX_SEED, level=lv)
        if best is None or st["ms_total"] < best["ms_total"]:
            best = st
    print("z%d ratio %.2f  device %.3e calls/s (%.2f ms, k_auto/k_lz %.2f ms = %.0f GB/s text)  first call %.2f s" % (
        lv, best["text_bytes"] / best["bgzf_bytes"], best["calls"] / (best["ms_total"] * 1e-3), best["ms_total"], best["ms_auto"],
        best["auto_text_bytes"] / (best["ms_auto"] * 1e-3) / 1e9, t_first), flush=True)
