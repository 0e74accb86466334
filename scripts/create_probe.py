"""Ad-hoc: cost of engine creation / destruction and of the first generate (allocation) vs the second."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
t0 = time.perf_counter()
from dna_factory_b200 import _native
import bench
t1 = time.perf_counter(); print("imports %.2f s" % (t1 - t0))
for i in range(3):
    t0 = time.perf_counter(); e = _native.Engine(0); t1 = time.perf_counter(); e.close(); t2 = time.perf_counter()
    print("create %.3f s, destroy %.3f s" % (t1 - t0, t2 - t1))
R = 65536
sex, ctl, table, orow, osamp = bench.synth_population(R, 0, window=R)
e = _native.Engine(0)
t0 = time.perf_counter(); e.set_samples(sex, ctl); e.set_snps(**table.device_arrays()); e.set_overrides(orow, osamp); t1 = time.perf_counter()
print("configure %.3f s" % (t1 - t0))
for i in range(3):
    n = [0]
    def w(b): n[0] += len(b)
    t0 = time.perf_counter(); st = e.generate_stream(0, R, 1, w, level=2); t1 = time.perf_counter()
    print("generate_stream %d: %.3f s (%d MB)" % (i, t1 - t0, n[0] >> 20))
import tempfile
with tempfile.NamedTemporaryFile() as f:
    for i in range(3):
        f.seek(0); t0 = time.perf_counter(); st = e.generate_fd(0, R, 1, f.fileno(), level=2); t1 = time.perf_counter()
        print("generate_fd %d: %.3f s (%d MB)" % (i, t1 - t0, st["bgzf_bytes"] >> 20))
t0 = time.perf_counter(); e.close(); print("destroy %.3f s" % (time.perf_counter() - t0))
