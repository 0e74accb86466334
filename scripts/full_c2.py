"""the device pipeline over the WHOLE SNP range of config C2 (and C4's shape per GPU) into a
discard-after-checksum sink: wall time of one dnaf_generate_device call."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from dna_factory_b200 import _native
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
S = int(sys.argv[2]) if len(sys.argv) > 2 else 5_000_000
bench.N_CASES = N // 2; bench.N_CONTROLS = N - N // 2
t0 = time.perf_counter()
sex, ctl, table, orow, osamp = bench.synth_population(S, 0, window=S)
arrays = table.device_arrays()
t1 = time.perf_counter()
eng = _native.Engine(0)
eng.set_samples(sex, ctl); eng.set_snps(**arrays); eng.set_overrides(orow, osamp)
eng.generate_device(0, min(S, 40000), 1, level=2)      # tables, allocations
t2 = time.perf_counter()
st = eng.generate_device(0, S, bench.PHILOX_SEED, level=2)
t3 = time.perf_counter()
print("N %d x S %d: host table %.1f s, configure+first pass %.1f s, generate_device %.3f s wall = %.3e calls/s "
      "(%.1f GB text, %.2f GB BGZF, ratio %.2f, %d blocks, crc xor %08x; event sums: fused %.0f ms, compaction %.0f ms)" % (
          N, S, t1 - t0, t2 - t1, t3 - t2, st["calls"] / (t3 - t2), st["text_bytes"] / 1e9, st["bgzf_bytes"] / 1e9,
          st["text_bytes"] / st["bgzf_bytes"], st["bgzf_blocks"], st["crc_xor"], st["ms_fused"], st["ms_deflate"]))
