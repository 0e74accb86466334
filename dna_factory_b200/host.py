"""Host-side flattening of the reference's objects into the arrays the C ABI takes.

Everything here is the string / dict work the reference does in Python around its row loop; none of it
is arithmetic on genotypes.  Reference lines mirrored:
  is_haploid(chromo, is_male)                      common/snp.py:102-109   -> chromosome class
  SNPTuples.pick_allele_index  `cum >= u`          pop_factory.py:92-95    -> integer thresholds
  row lead "%s\\t%i\\trs%s\\t%s\\t%s\\t40\\tPASS\\t.\\tGT\\t"    pop_factory.py:503-507
  SNPTuples.alt_alleles / ref_allele_tuple         pop_factory.py:104-116
  `snp.id not in sample.deleterious_snps`          pop_factory.py:485      -> override pairs
"""
import math

import numpy as np

from ._native import CLASS_AUTO, CLASS_MT, CLASS_X, CLASS_Y, KMAX

_CLASS = {"X": CLASS_X, "Y": CLASS_Y, "MT": CLASS_MT}
TWO32 = 4294967296.0


def chrom_class(chromosome):
    """Only 'X', 'Y' and 'MT' change ploidy (common/snp.py:109); every other label is diploid."""
    return _CLASS.get(chromosome, CLASS_AUTO)


def threshold(cum):
    """T with (cum >= U * 2**-32) <=> (U <= T) for every 32-bit U; cum * 2**32 is exact in float64."""
    if not (cum >= 0.0):
        raise ValueError("cumulative allele probability %r is negative or NaN" % (cum,))
    if cum >= 1.0:
        return 0xFFFFFFFF
    return min(int(math.floor(cum * TWO32)), 0xFFFFFFFF)


def alt_alleles(tuples):
    if len(tuples) == 1:
        return tuples[0][0]
    if len(tuples) == 2:
        return tuples[1][0]
    return ",".join(t[0] for t in tuples[1:])


def row_prefix(snp):
    return "%s\t%i\trs%s\t%s\t%s\t40\tPASS\t.\tGT\t" % (snp.chromosome, snp.position, snp.id, snp.tuples[0][0],
                                                          alt_alleles(snp.tuples))


def flatten_snps(snps):
    """list of SNPTuples-shaped objects -> dict(chrom_class, n_alleles, thresholds, prefix_bytes, prefix_off)."""
    S = len(snps)
    cls = np.zeros(S, dtype=np.uint8)
    k = np.zeros(S, dtype=np.uint8)
    thr = np.full((max(S, 1), KMAX), 0xFFFFFFFF, dtype=np.uint32)
    off = np.zeros(S + 1, dtype=np.uint64)
    parts = []
    pos = 0
    for r, s in enumerate(snps):
        n = len(s.tuples)
        if n < 1 or n > KMAX:
            raise ValueError("SNP %s has %d alleles; the device path handles 1..%d" % (s.id, n, KMAX))
        cls[r] = chrom_class(s.chromosome)
        k[r] = n
        for j, t in enumerate(s.tuples):
            thr[r, j] = threshold(t[1])
        b = row_prefix(s).encode("latin-1")
        parts.append(b)
        pos += len(b)
        off[r + 1] = pos
    prefix = np.frombuffer(b"".join(parts) + b"\0", dtype=np.uint8)
    return dict(chrom_class=cls, n_alleles=k, thresholds=thr[:S] if S else thr[:0], prefix_bytes=prefix,
                prefix_off=off)


def flatten_samples(fam_data):
    n = len(fam_data)
    sex = np.fromiter((s.sex for s in fam_data), dtype=np.int64, count=n)
    sex = np.where(sex == 1, 1, 2).astype(np.uint8)        # SampleInfo.is_male(): sex == 1
    ctl = np.fromiter((1 if s.is_control else 0 for s in fam_data), dtype=np.uint8, count=n)
    return sex, ctl


def override_pairs(fam_data, snps, row_base=0):
    """(global row, sample) pairs, sorted by row, where the reference forces the minor allele.

    Membership is tested exactly like the reference does (`snp.id in sample.deleterious_snps`), so a
    deleterious.json replay -- whose keys are strings while snp.id is an int -- yields no overrides,
    as it does upstream.
    """
    cases = [(i, s.deleterious_snps) for i, s in enumerate(fam_data)
             if not s.is_control and s.deleterious_snps]
    rows, samples = [], []
    if cases:
        keys = set()
        for _, d in cases:
            keys.update(d.keys())
        for r, snp in enumerate(snps):
            try:
                hit = snp.id in keys
            except TypeError:
                hit = False
            if hit:
                for i, d in cases:
                    if snp.id in d:
                        rows.append(row_base + r)
                        samples.append(i)
    return np.asarray(rows, dtype=np.uint64), np.asarray(samples, dtype=np.uint32)


def override_pairs_table(fam_data, table, row_base=0):
    """override_pairs for a SnpTable (integer ids): same membership semantics, no per-SNP Python objects.
    A key matches a row when `row_id in {key: ...}` would, i.e. when the key is a number equal to the id
    (string keys of a deleterious.json replay never equal an int id)."""
    ids = table.ids
    order = np.argsort(ids, kind="stable")
    sorted_ids = ids[order]
    rows, samples = [], []
    for i, s in enumerate(fam_data):
        if s.is_control or not s.deleterious_snps:
            continue
        for key in s.deleterious_snps.keys():
            if isinstance(key, bool) or not isinstance(key, (int, float, np.integer, np.floating)) or key != int(key):
                continue
            lo = np.searchsorted(sorted_ids, int(key), side="left")
            hi = np.searchsorted(sorted_ids, int(key), side="right")
            for r in order[lo:hi]:
                rows.append(row_base + int(r))
                samples.append(i)
    rows = np.asarray(rows, dtype=np.uint64)
    samples = np.asarray(samples, dtype=np.uint32)
    o = np.lexsort((samples, rows))
    return rows[o], samples[o]


def configure(engine, fam_data, snps):
    """Load a population (samples, sorted SNP list, overrides) into an Engine."""
    sex, ctl = flatten_samples(fam_data)
    engine.set_samples(sex, ctl)
    engine.set_snps(**flatten_snps(snps))
    rows, samples = override_pairs(fam_data, snps)
    engine.set_overrides(rows, samples)
    return engine
