"""Allele-frequency chi-square of the hot path's draws (BASELINE.json north_star: "allele-frequency chi-square checks
are additionally reported for the native RNG mode").

The reference's own stream is statistically defective: its forked workers share one numpy RNG state, so rows of the
same stripe replay the same uniforms and their minor-allele sets are nested (pop_factory.py:235,429-434,
477).  The counter-based stream here keys every (row, allele slot) separately, so the checks are against the MODEL
the reference samples from -- each allele is minor with probability maf, independently (pop_factory.py:477-494) --
not against reference output:
  1. per MAF bin (the 99 bins of snp_freq_cdf.csv): minor-allele counts of many rows against Binomial(2N, maf)
  2. rows that share a MAF are independent: joint minor counts of row pairs against 2N * maf^2
  3. the two allele slots of a sample are independent: joint counts against N * maf^2
The text comes out of the fused kernels (k_auto / k_lz) through the C ABI and is inflated by zlib.
Used by tests/test_gpu_stats.py and reported by bench.py (`allele_chi_square`).
"""
from types import SimpleNamespace

import numpy as np

from . import _native, host
from .maf_cdf import MAF_CDF


def chi2_sf(x, k):
    """Upper tail of the chi-square distribution with k degrees of freedom: Q(k/2, x/2), the regularised incomplete
    gamma function by its series / continued fraction (Numerical Recipes 6.2).  Pure Python: importing scipy.stats on
    a box with a cold page cache costs minutes, this costs microseconds."""
    import math
    a, x = 0.5 * k, 0.5 * x
    if x <= 0:
        return 1.0
    gln = math.lgamma(a)
    if x < a + 1.0:                       # series for P, return 1 - P
        ap, term, total = a, 1.0 / a, 1.0 / a
        for _ in range(100000):
            ap += 1.0
            term *= x / ap
            total += term
            if abs(term) < abs(total) * 1e-15:
                break
        return max(0.0, 1.0 - total * math.exp(-x + a * math.log(x) - gln))
    b = x + 1.0 - a                       # continued fraction for Q (modified Lentz)
    c = 1.0 / 1e-300
    d = 1.0 / b
    h = d
    for i in range(1, 100000):
        an = -i * (i - a)
        b += 2.0
        d = an * d + b
        d = 1e-300 if abs(d) < 1e-300 else d
        c = b + an / c
        c = 1e-300 if abs(c) < 1e-300 else c
        d = 1.0 / d
        delta = d * c
        h *= delta
        if abs(delta - 1.0) < 1e-15:
            break
    return min(1.0, math.exp(-x + a * math.log(x) - gln) * h)


def inflate_bgzf(blob):
    """Text of a BGZF stream, block by block through the BSIZE field (gzip.decompress() re-copies the rest of the
    stream after every member: minutes for the ten thousand blocks of a 400 MB text)."""
    import zlib
    view, out, o = memoryview(blob), [], 0
    while o < len(view):
        if view[o] != 0x1f or view[o + 1] != 0x8b or view[o + 12] != 0x42 or view[o + 13] != 0x43:
            raise ValueError("not a BGZF block at offset %d" % o)
        size = (view[o + 16] | (view[o + 17] << 8)) + 1
        text = zlib.decompress(view[o + 18:o + size - 8], -15)
        crc = int.from_bytes(view[o + size - 8:o + size - 4], "little")
        if zlib.crc32(text) != crc or len(text) != int.from_bytes(view[o + size - 4:o + size], "little"):
            raise ValueError("BGZF block at offset %d: CRC32 / ISIZE mismatch" % o)
        out.append(text)
        o += size
    return b"".join(out)


def allele_matrix(level, n_samples, rows_per_bin, seed, device=0):
    """[bins][rows_per_bin][2N] uint8 minor-allele indicators of autosome rows, one MAF bin after the other."""
    mafs = [m for m, _ in MAF_CDF]
    snps = [SimpleNamespace(id=1 + b * rows_per_bin + r, chromosome="1", position=1000 + b * rows_per_bin + r,
                            tuples=[("A", 1 - maf), ("C", 1.0)]) for b, maf in enumerate(mafs) for r in range(rows_per_bin)]
    import os, sys, time
    t0 = time.perf_counter()

    def lap(what):
        if os.environ.get("DNAF_STATS_TRACE"):
            print("[allele_stats] %-28s %.2f s" % (what, time.perf_counter() - t0), file=sys.stderr, flush=True)

    sex = (1 + (np.arange(n_samples) & 1)).astype(np.uint8)
    ctl = np.ones(n_samples, np.uint8)
    with _native.Engine(device) as eng:
        eng.set_samples(sex, ctl)
        eng.set_snps(**host.flatten_snps(snps))
        eng.set_overrides(np.zeros(0, np.uint64), np.zeros(0, np.uint32))
        lap("engine configured")
        blob, st = eng.generate(0, len(snps), seed, level=level)
        lap("generated")
    if not (st["ms_fused"] > 0 and st["calls"] == len(snps) * n_samples):
        raise RuntimeError("the fused kernels did not take these rows")
    text = inflate_bgzf(blob)
    lap("inflated")
    lines = text.split(b"\n")[:-1]
    if len(lines) != len(snps):
        raise RuntimeError("row count of the inflated text is wrong")
    out = np.empty((len(mafs), rows_per_bin, 2 * n_samples), dtype=np.uint8)
    for i, ln in enumerate(lines):
        body = np.frombuffer(ln[len(ln) - (4 * n_samples - 1):] + b"\t", dtype=np.uint8).reshape(n_samples, 4)
        out[i // rows_per_bin, i % rows_per_bin] = body[:, (0, 2)].reshape(-1) - 48
    lap("parsed")
    return np.array(mafs), out


def chi_square_report(level=2, n_samples=20000, rows_per_bin=51, seed=0x5EED000000000001, device=0):
    """>= 1e8 calls by default: 99 bins x 51 rows x 20000 samples.  Returns a dict of statistics."""
    mafs, a = allele_matrix(level, n_samples, rows_per_bin, seed, device)
    n_all = a.shape[2]
    counts = a.sum(axis=2, dtype=np.int64)                          # [bins][rows]
    exp = n_all * mafs[:, None]
    var = n_all * (mafs * (1 - mafs))[:, None]
    z2 = (counts - exp) ** 2 / var
    chi_bins = z2.sum(axis=1)                                       # ~ chi2(rows_per_bin) per bin
    p_bins = np.array([chi2_sf(float(c), rows_per_bin) for c in chi_bins])
    chi_total = float(z2.sum())
    dof = z2.size
    half = rows_per_bin // 2
    # independence of rows that share a MAF: joint minor counts of consecutive row pairs
    joint = (a[:, 0:2 * half:2] & a[:, 1:2 * half:2]).sum(axis=2, dtype=np.int64)
    pj = (mafs ** 2)[:, None]
    zj2 = (joint - n_all * pj) ** 2 / (n_all * pj * (1 - pj))
    # a sample's two allele slots are independent too (slot 2i vs 2i+1)
    within = (a[:, :, 0::2] & a[:, :, 1::2]).sum(axis=2, dtype=np.int64)
    zw2 = (within - (n_all // 2) * pj) ** 2 / ((n_all // 2) * pj * (1 - pj))
    return {"calls": int(a.shape[0] * a.shape[1] * n_samples), "level": level, "bins": int(len(mafs)), "rows_per_bin": rows_per_bin,
            "chi2": chi_total, "dof": int(dof), "p_value": chi2_sf(chi_total, dof),
            "min_bin_p": float(p_bins.min()), "worst_bin_maf": float(mafs[int(p_bins.argmin())]),
            "pair_chi2": float(zj2.sum()), "pair_dof": int(zj2.size), "pair_p_value": chi2_sf(float(zj2.sum()), zj2.size),
            "slot_chi2": float(zw2.sum()), "slot_dof": int(zw2.size), "slot_p_value": chi2_sf(float(zw2.sum()), zw2.size),
            "max_abs_freq_error": float(np.abs(counts.sum(axis=1) / (rows_per_bin * n_all) - mafs).max())}
