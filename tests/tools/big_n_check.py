import os, sys; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, time
from dna_factory_b200 import _native, host
from oracle import oracle
from tests.cases import synth_case
for n in (300000, 1000003):
    case = synth_case(n, 4, seed=5, chroms=['1','X','Y','2'], n_del=2)
    t0=time.time(); want,_ = oracle.rows(case.samples, case.snps, case.seed, 0, n_threads=8); t1=time.time()
    eng=_native.Engine(0); host.configure(eng, case.samples, case.snps)
    blob, st = eng.generate(0, 4, case.seed, level=2)
    ok = oracle.bgzf_decompress(blob)[0] == want
    print(n, ok, st["bgzf_blocks"], "ratio %.2f"%(st["text_bytes"]/st["bgzf_bytes"]), "oracle %.1fs"%(t1-t0))
    eng.close()
