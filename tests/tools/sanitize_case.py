"""Small multi-kernel case for compute-sanitizer: autosome + X + Y rows, overrides, SNP selection, fd sink."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from dna_factory_b200 import _native, host, snp
from oracle import oracle
from tests.cases import synth_case
for n, s in ((2100, 9), (130, 12)):
    case = synth_case(n, s, seed=3, chroms=['1', 'X', 'Y', 'MT', '2'], n_del=5, exotic=True)
    want, _ = oracle.rows(case.samples, case.snps, case.seed, 0)
    eng = _native.Engine(0)
    host.configure(eng, case.samples, case.snps)
    eng.set_chunk_bytes(1 << 16)
    blob, st = eng.generate(0, s, case.seed, level=2)
    assert oracle.bgzf_decompress(blob)[0] == want
    with tempfile.TemporaryFile() as f:
        eng.generate_fd(0, s, case.seed, f.fileno(), level=2)
    eng.genotypes(0, s, case.seed); eng.text(0, s, case.seed)
    fac = snp.SnpFactory.init_from_cdf_file()
    fac.random_snp_table_device(eng, 5000, 7, min_maf=0.01)
    eng.close()
print("sanitize case ok")
