"""CPU tests of the LZ tiers' (-z 3..9) span grammar, static code tables and dynamic-block header: the host twin of
k_lz (same lz_span_tokens function, same table builder) encodes synthetic autosome segments, zlib inflates them.
No GPU involved; the device path is checked against the oracle in tests/test_gpu_parity.py."""
import zlib

import numpy as np
import pytest

from dna_factory_b200 import _native


def _segment(n_cells, p, seed, first_byte_tab, ends_row, forced=None):
    rs = np.random.RandomState(seed)
    alleles = (rs.rand(2 * n_cells) < p).astype(np.uint8)
    if forced is not None:
        alleles[forced] = 1
    bits = np.zeros((2 * n_cells + 31) // 32 + 1, dtype=np.uint32)
    idx = np.nonzero(alleles)[0]
    np.bitwise_or.at(bits, idx >> 5, (np.uint32(1) << (idx & 31).astype(np.uint32)))
    t = np.empty(4 * n_cells, np.uint8)
    t[0::4] = alleles[0::2] + 48
    t[1::4] = ord("/")
    t[2::4] = alleles[1::2] + 48
    t[3::4] = 9
    body = t.tobytes()
    body = body[:-1] + (b"\n" if ends_row else b"")     # a segment does not own its last separator unless it ends the row
    return bits, (b"\t" if first_byte_tab else b"") + body


@pytest.mark.parametrize("level", [3, 4, 5, 6, 7, 8, 9])
@pytest.mark.parametrize("p", [0.005, 0.03, 0.11, 0.2, 0.35, 0.495])
def test_lz_block_inflates_to_the_text(level, p):
    for n_cells, with_prefix, ends_row, seed in ((10048, True, False, 1), (9952, False, True, 2), (777, True, True, 3),
                                                 (64, False, False, 4), (1, True, True, 5), (16256, False, True, 6)):
        prefix = b"12\t34567\trs89\tA\tC\t40\tPASS\t.\tGT\t" if with_prefix else b""
        bits, body = _segment(n_cells, p, 100 * level + seed, not with_prefix, ends_row)
        enc = _native.debug_lz_block(p, level, bits, n_cells, prefix, ends_row)
        assert zlib.decompress(enc, -15) == prefix + body
        assert len(enc) < len(prefix + body) // 2 + 64


def test_lz_block_handles_patterns_the_tables_never_expect():
    """Forced-minor runs (pop_factory.py:495-499) on a rare-MAF table: every symbol must still have a code."""
    n_cells = 9000
    forced = np.arange(4000, 9000)           # thousands of 1/1 cells in a row on a MAF 0.005 table
    bits, body = _segment(n_cells, 0.005, 7, True, False, forced)
    for level in (3, 4, 6, 9):
        enc = _native.debug_lz_block(0.005, level, bits, n_cells, b"", False)
        assert zlib.decompress(enc, -15) == body


def test_lz_levels_are_monotone_on_the_reference_maf_mix():
    """Deeper tiers must not compress worse (pop_factory.py:403 hands -z to the writer; BASELINE config 5)."""
    sizes = {}
    for level in (3, 4, 5, 6, 7, 8, 9):
        tot = 0
        for i, p in enumerate((0.01, 0.03, 0.08, 0.15, 0.3, 0.45)):
            bits, body = _segment(10048, p, 40 + i, True, False)
            tot += len(_native.debug_lz_block(p, level, bits, 10048, b"", False))
        sizes[level] = tot
    assert all(sizes[a] >= sizes[b] for a, b in zip((3, 4, 5, 6, 7, 8), (4, 5, 6, 7, 8, 9))), sizes
