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

# dense forced-minor patterns on rare rows: k_auto's staging-overflow re-emit (atomicOr into the slot words) and the
# stored-block fallback, k_x's the same (tests/test_gpu_parity.py::test_auto_kernel_dense_overrides_on_rare_rows)
from types import SimpleNamespace
from tests.cases import Snp, Sample
n, stride = 8200, 3
snps = [Snp(id=i + 1, chromosome='1', position=1000 * (i + 1), tuples=[("A", 1 - maf), ("C", 1.0)])
        for i, maf in enumerate([0.005, 0.01, 0.02, 0.25, 0.495, 0.005])]
snps += [Snp(id=7, chromosome='X', position=5, tuples=[("A", 0.995), ("C", 1.0)]),
         Snp(id=8, chromosome='X', position=6, tuples=[("A", 0.6), ("C", 1.0)]),
         Snp(id=9, chromosome='X', position=7, tuples=[("G", 1.0)])]
samples = []
for i in range(n):
    ctl = i < n // 3
    d = None
    if not ctl:
        d = {s.id: 0.5 for s in snps[:5] + snps[6:]} if (i % stride == 0) else {}
        if n // 2 <= i < n // 2 + 70:
            d[6] = 0.5
    samples.append(Sample(family_id=i + 1, person_id=100001 + i, father_id=0, mother_id=0, sex=1 + (i & 1),
                          is_control=ctl, deleterious_snps=d))
want, _ = oracle.rows(samples, snps, 0xD15EA5E, 0, n_threads=4)
eng = _native.Engine(0)
host.configure(eng, samples, snps)
for level in (2, 6, 9):
    blob, st = eng.generate(0, len(snps), 0xD15EA5E, level=level)
    assert st["ms_fused"] > 0 and oracle.bgzf_decompress(blob)[0] == want
eng.close()
print("sanitize dense-override case ok")
