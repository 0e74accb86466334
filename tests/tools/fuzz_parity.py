"""Randomised parity soak: populations of random size / sex ratio / MAF mix / override density / pass size through the
CUDA path (C ABI) against the CPU oracle.  Usage: python tests/tools/fuzz_parity.py [cases] [seed]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from types import SimpleNamespace
from dna_factory_b200 import _native, host
from oracle import oracle
from tests.cases import Snp, Sample

cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rs = np.random.RandomState(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
t0 = time.time()
bad = 0
for it in range(cases):
    n = int(rs.choice([rs.randint(1, 1100), rs.randint(1024, 9000), rs.randint(9000, 70000)]))
    s = int(rs.randint(1, 14))
    odds = float(rs.choice([0.0, 1.0, 0.5, rs.rand()]))
    chroms = [str(c) for c in rs.choice(['1', '2', '22', 'X', 'X', 'Y', 'MT'], size=s)]
    snps = []
    for i in range(s):
        kind = rs.randint(8)
        if kind == 0:
            tup = [("A", 1.0)]
        elif kind == 1:
            tup = [("A", 0.5), ("C", 0.75), ("G", 1.0)]
        else:
            maf = float(rs.choice([0.005, 0.01, 0.495, 0.25, 1e-6, rs.rand() * 0.5]))
            tup = [("T", 1 - maf), ("G", 1.0)]
        snps.append(Snp(id=i + 1, chromosome=chroms[i], position=int(rs.randint(0, 10 ** 8)), tuples=tup))
    snps.sort(key=lambda x: (x.chromosome, x.position))
    dens = float(rs.choice([0.0, 0.01, 0.5, 1.0]))
    samples = []
    for i in range(n):
        ctl = i < n // 2
        d = None if ctl else {sn.id: 1.0 for sn in snps if rs.rand() < dens}
        samples.append(Sample(family_id=i + 1, person_id=100001 + i, father_id=0, mother_id=0,
                              sex=1 if rs.rand() < odds else 2, is_control=ctl, deleterious_snps=d))
    seed = int(rs.randint(1, 2 ** 31))
    row_base = int(rs.choice([0, 5, 2 ** 32 - 3]))
    want, _ = oracle.rows(samples, snps, seed, row_base, n_threads=8)
    eng = _native.Engine(0)
    host.configure(eng, samples, snps)
    eng.set_row_base(row_base)
    eng.set_chunk_bytes(int(rs.choice([4096, 1 << 16, 1 << 20, 1 << 30])))
    blob, st = eng.generate(0, s, seed, level=int(rs.randint(1, 10)))
    try:
        text, blocks, _ = oracle.bgzf_decompress(blob)
        ok = text == want and blocks == st["bgzf_blocks"] and st["calls"] == n * s
    except Exception as e:
        ok = False
        print("   exception:", e)
    if not ok:
        bad += 1
        print("MISMATCH case %d: n=%d s=%d odds=%.2f dens=%.2f chroms=%s seed=%d row_base=%d" % (it, n, s, odds, dens, chroms, seed, row_base))
    eng.close()
print("%d cases, %d mismatches, %.1f s" % (cases, bad, time.time() - t0))
sys.exit(1 if bad else 0)
