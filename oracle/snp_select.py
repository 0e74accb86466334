"""numpy restatement of the SNP selection (TEST INFRASTRUCTURE ONLY): SnpFactory.random_snp_tuples
(pop_factory.py:160-193) driven by the counter-based replay stream, and the sort of pop_factory.py:245.

Stream spec (csrc/k_select.cuh): for draw n
  A = philox4x32_10(ctr=(n_lo, n_hi, 0x534E5000, 0xFFFFFFFF), key=seed): u_chrom, u_maf, u_pos, u_ref = A * 2**-32
  B = philox4x32_10(ctr=(n_lo, n_hi, 0x534E5001, 0xFFFFFFFF), key=seed): u_alt = B[0] * 2**-32
"""
import numpy as np

from . import philox_np

TAG_A, TAG_B = 0x534E5000, 0x534E5001


def uniforms(seed, size):
    """-> dict of float64 arrays chrom, maf, pos, ref, alt (one uniform per draw each)."""
    n = np.arange(size, dtype=np.uint64)
    ctr = np.zeros((size, 4), dtype=np.uint64)
    ctr[:, 0] = n & np.uint64(0xFFFFFFFF)
    ctr[:, 1] = n >> np.uint64(32)
    ctr[:, 3] = 0xFFFFFFFF
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    ctr[:, 2] = TAG_A
    a = philox_np.philox4x32_10(ctr, key).astype(np.float64) * 2.0 ** -32
    ctr[:, 2] = TAG_B
    b = philox_np.philox4x32_10(ctr, key).astype(np.float64) * 2.0 ** -32
    return dict(chrom=a[:, 0], maf=a[:, 1], pos=a[:, 2], ref=a[:, 3], alt=b[:, 0])


def select(seed, size, chrom_cdf, chrom_max_pos, chrom_rank, maf_cdf, sort=True):
    """Columns (order, chrom_idx, maf_bin, position, ref, alt) as csrc/k_select.cuh produces them."""
    u = uniforms(seed, size)
    chrom = np.searchsorted(chrom_cdf, u["chrom"], side="right")           # numpy.random.choice(p=...)
    maf = np.searchsorted(maf_cdf, u["maf"], side="right")
    pos = (u["pos"] * np.asarray(chrom_max_pos)[chrom]).astype(np.int64)    # int(u * max)
    ref_idx = (u["ref"] * 4.0).astype(np.int64)                            # choice without p = randint(4)
    pick = (u["alt"] * 3.0).astype(np.int64)
    alt_idx = pick + (pick >= ref_idx)                                      # remaining_nt.remove(ref); random.choice
    codes = np.frombuffer(b"ATCG", dtype=np.uint8)
    order = np.arange(size)
    if sort:
        order = np.lexsort((pos, np.asarray(chrom_rank)[chrom]))           # stable: ties keep draw order
    return dict(order=order.astype(np.uint32), chrom_idx=chrom[order].astype(np.uint8), maf_bin=maf[order].astype(np.uint8),
                position=pos[order].astype(np.uint32), ref=codes[ref_idx][order], alt=codes[alt_idx][order])
