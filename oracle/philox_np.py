"""numpy restatement of the replay uniform stream (TEST INFRASTRUCTURE ONLY).

Spec, identical to oracle/dnaf_oracle.c:
  W(row, g, b) = philox4x32_10(ctr=(g, b>>2, row_lo, row_hi), key=(seed_lo, seed_hi))[b & 3]
  bit (31-b) of U_j = bit (j & 31) of W(row, j>>5, b);   u_j = U_j * 2**-32
``uniforms(seed, row, n)`` is what the patched ``numpy.random.rand(n)`` returns to the
reference's worker loop (pop_factory.py:477) for the SNP whose global sorted row index is `row`.
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr: (..., 4) uint32 array, key: (2,) ints -> (..., 4) uint32."""
    c = np.asarray(ctr, dtype=np.uint64)
    c0, c1, c2, c3 = c[..., 0], c[..., 1], c[..., 2], c[..., 3]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, p1 & MASK, n2, p0 & MASK
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def uniform_bits(seed, row, n_alleles):
    """32-bit integer uniforms U_j, j < n_alleles, for one SNP row."""
    groups = (n_alleles + 31) // 32
    if groups == 0:
        return np.zeros(0, dtype=np.uint32)
    g = np.arange(groups, dtype=np.uint64)
    q = np.arange(8, dtype=np.uint64)
    ctr = np.zeros((groups, 8, 4), dtype=np.uint64)
    ctr[..., 0] = g[:, None]
    ctr[..., 1] = q[None, :]
    ctr[..., 2] = row & 0xFFFFFFFF
    ctr[..., 3] = (row >> 32) & 0xFFFFFFFF
    w = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)).reshape(groups, 32)  # [g][b]
    lanes = np.arange(32, dtype=np.uint32)
    bits = (w[:, :, None] >> lanes[None, None, :]) & np.uint32(1)  # [g][b][lane]
    weights = (np.uint64(1) << (np.uint64(31) - np.arange(32, dtype=np.uint64)))  # b -> 2^(31-b)
    U = (bits.astype(np.uint64) * weights[None, :, None]).sum(axis=1).astype(np.uint32)  # [g][lane]
    return U.reshape(-1)[:n_alleles]


def uniforms(seed, row, n_alleles):
    return uniform_bits(seed, row, n_alleles).astype(np.float64) * 2.0 ** -32
