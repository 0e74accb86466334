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
    (string keys of a deleterious.json replay never equal an int id).  One pass collects the (sample, key) pairs,
    the id lookup is a single vectorised searchsorted."""
    def numeric(key):
        return (not isinstance(key, bool)) and isinstance(key, (int, float, np.integer, np.floating)) and key == int(key)

    pairs = [(i, int(key)) for i, s in enumerate(fam_data) if not s.is_control and s.deleterious_snps
             for key in s.deleterious_snps.keys() if numeric(key)]
    if not pairs:
        return np.zeros(0, np.uint64), np.zeros(0, np.uint32)
    samp = np.fromiter((p[0] for p in pairs), dtype=np.int64, count=len(pairs))
    keys = np.fromiter((p[1] for p in pairs), dtype=np.int64, count=len(pairs))
    ids = np.asarray(table.ids, dtype=np.int64)
    order = np.argsort(ids, kind="stable")
    sorted_ids = ids[order]
    lo = np.searchsorted(sorted_ids, keys, side="left")
    hi = np.searchsorted(sorted_ids, keys, side="right")
    cnt = hi - lo                                            # rows carrying that id (1 for SnpFactory output)
    tot = int(cnt.sum())
    if tot == 0:
        return np.zeros(0, np.uint64), np.zeros(0, np.uint32)
    rep = np.repeat(np.arange(len(pairs)), cnt)
    within = np.arange(tot) - np.repeat(np.cumsum(cnt) - cnt, cnt)
    rows = (order[lo[rep] + within] + row_base).astype(np.uint64)
    samples = samp[rep].astype(np.uint32)
    o = np.lexsort((samples, rows))
    return rows[o], samples[o]


def slice_snps(arrays, lo, hi):
    """Rows [lo, hi) of flatten_snps() / SnpTable.device_arrays() output, as their own dnaf_set_snps arguments."""
    p0, p1 = int(arrays["prefix_off"][lo]), int(arrays["prefix_off"][hi])
    return dict(chrom_class=arrays["chrom_class"][lo:hi], n_alleles=arrays["n_alleles"][lo:hi],
                thresholds=arrays["thresholds"][lo:hi],
                prefix_bytes=np.concatenate([arrays["prefix_bytes"][p0:p1], np.zeros(1, np.uint8)]),
                prefix_off=arrays["prefix_off"][lo:hi + 1] - np.uint64(p0))


def slice_overrides(orow, osamp, lo, hi):
    """Override pairs of rows [lo, hi), rows made local to the slice."""
    a, b = np.searchsorted(orow, [lo, hi], side="left")
    return (orow[a:b] - np.uint64(lo)).astype(np.uint64), osamp[a:b]


def configure(engine, fam_data, snps):
    """Load a population (samples, sorted SNP list, overrides) into an Engine."""
    sex, ctl = flatten_samples(fam_data)
    engine.set_samples(sex, ctl)
    engine.set_snps(**flatten_snps(snps))
    rows, samples = override_pairs(fam_data, snps)
    engine.set_overrides(rows, samples)
    return engine
