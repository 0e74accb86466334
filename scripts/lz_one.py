"""Ad-hoc: one device-only call at a given level (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from dna_factory_b200 import _native
lv = int(sys.argv[1]); R = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
sex, ctl, table, orow, osamp = bench.synth_population(R, 0, window=R)
eng = _native.Engine(0)
eng.set_samples(sex, ctl); eng.set_snps(**table.device_arrays()); eng.set_overrides(orow, osamp)
for _ in range(3):
    st = eng.generate_device(0, R, bench.PHILOX_SEED, level=lv)
print("z%d ratio %.2f %.3e calls/s k %.2f ms" % (lv, st["text_bytes"] / st["bgzf_bytes"], st["calls"] / (st["ms_total"] * 1e-3), st["ms_auto"]))
