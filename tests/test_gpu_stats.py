"""GPU test of dna_factory_b200/allele_stats.py: allele-frequency chi-square of the hot path's own draws (BASELINE.json north_star: "allele-frequency chi-square
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
import numpy as np
import pytest

from dna_factory_b200.allele_stats import chi_square_report


@pytest.mark.gpu
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
    from dna_factory_b200.allele_stats import chi2_sf
    rs = np.random.RandomState(5)
    u = rs.rand(40000)
    a, b = (u > 1 - 0.2).astype(np.uint8), (u > 1 - 0.2).astype(np.uint8)    # same stripe, same uniforms (R7)
    joint = int((a & b).sum())
    pj = 0.04
    z2 = (joint - 40000 * pj) ** 2 / (40000 * pj * (1 - pj))
    assert chi2_sf(z2, 1) < 1e-12
