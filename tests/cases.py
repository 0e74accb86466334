"""Loads the committed golden fixtures (tests/golden/*.json + *.rows.gz) into reference-shaped objects."""
import gzip
import hashlib
import json
import os
from types import SimpleNamespace

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ROW_CASES = ["mixed64", "r8_strkeys", "n0", "wide", "all_male", "all_female"]


class Snp(SimpleNamespace):
    """Same attributes as the reference's SNPTuples (pop_factory.py:74-85)."""


class Sample(SimpleNamespace):
    """Same attributes as the reference's SampleInfo (pop_factory.py:47-60)."""

    def is_male(self):
        return self.sex == 1


def load_case(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        meta = json.load(f)
    with gzip.open(os.path.join(GOLDEN, name + ".rows.gz"), "rb") as f:
        text = f.read()
    assert len(text) == meta["text_len"]
    assert hashlib.sha256(text).hexdigest() == meta["text_sha256"]
    snps = [Snp(id=s["id"], chromosome=s["chromosome"], position=s["position"],
                tuples=[(t[0], t[1]) for t in s["tuples"]]) for s in meta["snps"]]
    samples = []
    for s in meta["samples"]:
        d = None
        if s["deleterious"] is not None:
            d = {(int(k) if kind == "int" else str(k)): w for k, kind, w in s["deleterious"]}
        samples.append(Sample(family_id=s["family_id"], person_id=s["person_id"], father_id=0, mother_id=0,
                              sex=s["sex"], is_control=s["is_control"], deleterious_snps=d))
    return SimpleNamespace(name=name, seed=meta["seed"], row_begin=meta["row_begin"], samples=samples, snps=snps,
                           text=text)


def synth_case(n_samples, n_snps, seed=1, male_odds=0.5, chroms=None, n_case_frac=0.5, n_del=3, exotic=False):
    """Seeded synthetic population shaped like SnpFactory output (biallelic, MAF grid) for oracle-vs-CUDA tests."""
    import numpy as np
    rs = np.random.RandomState(seed)
    chroms = chroms or ['1', '2', '7', '12', '22', 'X', 'Y']
    snps = []
    for i in range(n_snps):
        c = chroms[rs.randint(len(chroms))]
        maf = (1 + rs.randint(99)) * 0.005
        nts = ["A", "T", "C", "G"]
        rs.shuffle(nts)
        tup = [(nts[0], 1 - maf), (nts[1], 1.0)]
        if exotic and i % 5 == 1:
            tup = [(nts[0], 0.55), (nts[1], 0.8), (nts[2], 1.0)]
        elif exotic and i % 5 == 2:
            tup = [(nts[0], 0.4), (nts[1], 0.7), (nts[2], 0.9), (nts[3], 1.0)]
        elif exotic and i % 5 == 3:
            tup = [(nts[0], 1.0)]
        snps.append(Snp(id=i + 1, chromosome=c, position=int(rs.rand() * 5e7), tuples=tup))
    snps.sort(key=lambda x: (x.chromosome, x.position))
    n_ctl = int(n_samples * (1 - n_case_frac))
    samples = []
    ids = [s.id for s in snps]
    for i in range(n_samples):
        ctl = i < n_ctl
        d = None
        if not ctl:
            d = {int(k): 0.5 for k in rs.choice(ids, size=min(n_del, len(ids)), replace=False)}
        samples.append(Sample(family_id=i + 1, person_id=(100001 + i) if ctl else (500001 + i - n_ctl), father_id=0,
                              mother_id=0, sex=1 if rs.rand() <= male_odds else 2, is_control=ctl,
                              deleterious_snps=d))
    return SimpleNamespace(name="synth", seed=0xC0FFEE + seed, row_begin=0, samples=samples, snps=snps, text=None)
