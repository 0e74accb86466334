"""Allele-frequency chi-square of the hot path's own draws (BASELINE.json north_star: "allele-frequency chi-square
checks are additionally reported for the native RNG mode").

The reference's stream is statistically defective: its forked workers share one numpy RNG state, so rows of the same
stripe replay the same uniforms and their minor-allele sets are nested (pop_factory.py:235,429-434,477).
The counter-based stream here keys every (row, allele slot) separately, so the checks are against the MODEL the
reference samples from -- each allele is minor with probability maf, independently (pop_factory.py:477-494) -- not
against reference output:
  1. per MAF bin (the 99 bins of snp_freq_cdf.csv): minor-allele counts of many rows against Binomial(2N, maf)
  2. rows that share a MAF are independent: joint minor counts of row pairs against 2N * maf^2
The text comes out of the fused kernels (k_auto / k_lz) through the C ABI and is inflated by zlib.
"""
import gzip

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def allele_matrix(level, n_samples, rows_per_bin, seed):
    """[bins][rows_per_bin][2N] uint8 minor-allele indicators of autosome rows, one MAF bin after the other."""
    from dna_factory_b200 import _native, host
    from dna_factory_b200.maf_cdf import MAF_CDF
    from tests.cases import Snp, Sample
    mafs = [m for m, _ in MAF_CDF]
    snps = [Snp(id=1 + b * rows_per_bin + r, chromosome="1", position=1000 + b * rows_per_bin + r,
                tuples=[("A", 1 - maf), ("C", 1.0)]) for b, maf in enumerate(mafs) for r in range(rows_per_bin)]
    samples = [Sample(family_id=i + 1, person_id=100001 + i, father_id=0, mother_id=0, sex=1 + (i & 1), is_control=True,
                      deleterious_snps=None) for i in range(n_samples)]
    with _native.Engine(0) as eng:
        host.configure(eng, samples, snps)
        blob, st = eng.generate(0, len(snps), seed, level=level)
    assert st["ms_fused"] > 0 and st["calls"] == len(snps) * n_samples
    text = gzip.decompress(blob + _native.bgzf_eof())
    lines = text.split(b"\n")[:-1]
    assert len(lines) == len(snps)
    out = np.empty((len(mafs), rows_per_bin, 2 * n_samples), dtype=np.uint8)
    for i, ln in enumerate(lines):
        body = np.frombuffer(ln[len(ln) - (4 * n_samples - 1):] + b"\t", dtype=np.uint8).reshape(n_samples, 4)
        out[i // rows_per_bin, i % rows_per_bin] = body[:, (0, 2)].reshape(-1) - 48
    return np.array(mafs), out


def chi_square_report(level=2, n_samples=20000, rows_per_bin=51, seed=0x5EED000000000001):
    """>= 1e8 calls: 99 bins x 51 rows x 20000 samples.  Returns the statistics bench.py reports."""
    from scipy import stats
    mafs, a = allele_matrix(level, n_samples, rows_per_bin, seed)
    n_all = a.shape[2]
    counts = a.sum(axis=2, dtype=np.int64)                          # [bins][rows]
    exp = n_all * mafs[:, None]
    var = n_all * (mafs * (1 - mafs))[:, None]
    z2 = (counts - exp) ** 2 / var
    chi_bins = z2.sum(axis=1)                                       # ~ chi2(rows_per_bin) per bin
    p_bins = stats.chi2.sf(chi_bins, rows_per_bin)
    chi_total = float(z2.sum())
    dof = z2.size
    # independence of rows that share a MAF: joint minor counts of consecutive row pairs
    joint = (a[:, 0::2][:, : rows_per_bin // 2] & a[:, 1::2][:, : rows_per_bin // 2]).sum(axis=2, dtype=np.int64)
    pj = (mafs ** 2)[:, None]
    zj2 = (joint - n_all * pj) ** 2 / (n_all * pj * (1 - pj))
    # a row's two allele slots of a sample are independent too (slot 2i vs 2i+1)
    within = (a[:, :, 0::2] & a[:, :, 1::2]).sum(axis=2, dtype=np.int64)
    zw2 = (within - (n_all // 2) * pj) ** 2 / ((n_all // 2) * pj * (1 - pj))
    return {"calls": int(a.shape[0] * a.shape[1] * n_samples), "bins": int(len(mafs)), "rows_per_bin": rows_per_bin,
            "chi2": chi_total, "dof": int(dof), "p_value": float(stats.chi2.sf(chi_total, dof)),
            "min_bin_p": float(p_bins.min()), "worst_bin_maf": float(mafs[int(p_bins.argmin())]),
            "pair_chi2": float(zj2.sum()), "pair_dof": int(zj2.size), "pair_p_value": float(stats.chi2.sf(zj2.sum(), zj2.size)),
            "slot_chi2": float(zw2.sum()), "slot_dof": int(zw2.size), "slot_p_value": float(stats.chi2.sf(zw2.sum(), zw2.size)),
            "max_abs_freq_error": float(np.abs(counts.sum(axis=1) / (rows_per_bin * n_all) - mafs).max())}


@pytest.mark.parametrize("level", [2, 6])
def test_allele_frequencies_chi_square(level):
    r = chi_square_report(level=level)
    assert r["calls"] >= 10 ** 8
    # two-sided sanity: neither too far from the model nor suspiciously close to it
    assert 1e-4 < r["p_value"] < 1 - 1e-4, r
    assert r["min_bin_p"] > 1e-3 / r["bins"], r          # Bonferroni over the 99 bins
    assert 1e-4 < r["pair_p_value"] < 1 - 1e-4, r        # under the reference's R7 defect this statistic explodes
    assert 1e-4 < r["slot_p_value"] < 1 - 1e-4, r
    assert r["max_abs_freq_error"] < 2e-3, r


def test_chi_square_detects_the_reference_defect():
    """The pair statistic must be able to see nested minor sets: feed it two rows drawn from ONE uniform vector."""
    from scipy import stats
    rs = np.random.RandomState(5)
    u = rs.rand(40000)
    a, b = (u > 1 - 0.2).astype(np.uint8), (u > 1 - 0.2).astype(np.uint8)    # same stripe, same uniforms (R7)
    joint = int((a & b).sum())
    pj = 0.04
    z2 = (joint - 40000 * pj) ** 2 / (40000 * pj * (1 - pj))
    assert stats.chi2.sf(z2, 1) < 1e-12
