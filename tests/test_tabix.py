"""The .tbi index: builder against an independent reader of the format, on CPU.

The BGZF streams here are made with zlib (the test's own writer, random block cuts) -- the host-only exports
dnaf_bgzf_scan / the numpy builder are what is under test; the GPU run of the CLI with --tbi is in test_gpu_cli.py.
"""
import random
import struct
import zlib

import numpy as np
import pytest

from dna_factory_b200 import _native, tabix
from tests import tbi_reader

EOF_BLOCK = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def bgzf_block(text):
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = co.compress(text) + co.flush()
    return (b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(body) + 25) + body
            + struct.pack("<II", zlib.crc32(text), len(text)))


def bgzf_stream(text, cuts):
    """BGZF blocks holding text[cuts[i]:cuts[i+1]] (no EOF block)."""
    return b"".join(bgzf_block(text[a:b]) for a, b in zip(cuts[:-1], cuts[1:]))


def chrom_sorted_rows(rng, n, labels):
    """(label, pos) sorted the way pop_factory.py:245 sorts: chromosome STRING, then position."""
    rows = [(rng.choice(labels), rng.randrange(0, 3_000_000)) for _ in range(n)]
    rows.sort(key=lambda r: (r[0], r[1]))
    return rows


def make_case(seed, n_rows=400, n_samples=5, exact_boundaries=False):
    rng = random.Random(seed)
    header = b"##fileformat=VCFv4.1\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + \
        b"\t".join(b"%d" % i for i in range(n_samples)) + b"\n"
    labels = ["1", "10", "2", "X", "Y"]
    rows = chrom_sorted_rows(rng, n_rows, labels)
    lines = [b"%s\t%d\trs%d\tA\tC\t40\tPASS\t.\tGT\t" % (c.encode(), p, i) + b"\t".join([b"0/1"] * n_samples) + b"\n"
             for i, (c, p) in enumerate(rows)]
    body = b"".join(lines)
    row_off = np.concatenate(([0], np.cumsum([len(x) for x in lines]))).astype(np.uint64)
    # header in its own blocks (like BgzfSink.flush), body cut at random places
    hb = bgzf_stream(header, [0, len(header)])
    if exact_boundaries:
        cuts = sorted({0, len(body)} | {int(row_off[i]) for i in range(0, n_rows, 7)})
    else:
        cuts = sorted({0, len(body)} | {rng.randrange(1, len(body)) for _ in range(n_rows // 9)})
    bb = bgzf_stream(body, cuts)
    return header, hb, body, bb, rows, labels, row_off


def build(hb, bb, rows, labels, row_off):
    t = tabix.TabixBuilder()
    t.add_blocks(*_native.bgzf_scan(hb))
    ci = np.array([labels.index(c) for c, _ in rows])
    pos = np.array([p for _, p in rows])
    t.add_rows(labels, ci, pos, row_off)
    t.add_blocks(*_native.bgzf_scan(bb))
    return t.payload()


def brute(body, chrom, beg, end):
    out = []
    for line in body.splitlines(keepends=True):
        f = line.split(b"\t", 2)
        p0 = max(int(f[1]) - 1, 0)
        if f[0] == chrom.encode() and p0 < end and p0 + 1 > beg:
            out.append(line)
    return out


def test_bgzf_scan_block_table():
    text = bytes(random.Random(1).randrange(256) for _ in range(5000))
    blob = bgzf_stream(text, [0, 100, 100, 4000, 5000]) + EOF_BLOCK
    cs, us = _native.bgzf_scan(blob)
    assert us.tolist() == [100, 0, 3900, 1000, 0]
    assert int(cs.sum()) == len(blob) and cs[-1] == 28
    assert tbi_reader.bgzf_inflate_all(blob) == text
    cs0, us0 = _native.bgzf_scan(b"")
    assert len(cs0) == 0 and len(us0) == 0
    with pytest.raises(_native.DnafError):
        _native.bgzf_scan(blob[:-3])            # truncated last block
    with pytest.raises(_native.DnafError):
        _native.bgzf_scan(b"\x1f\x8b" + b"\0" * 40)   # gzip magic but no BC subfield


def test_virtual_offsets_boundaries():
    cs = np.array([50, 28, 70], np.uint32)
    us = np.array([100, 0, 200], np.uint32)
    v = tabix.virtual_offsets(np.array([0, 99, 100, 101, 299, 300], np.uint64), cs, us)
    # offset 100 is the start of the block after the empty one; 300 (end of text) the start of whatever follows
    assert v.tolist() == [0, 99, (78 << 16), (78 << 16) | 1, (78 << 16) | 199, 148 << 16]
    with pytest.raises(ValueError):
        tabix.virtual_offsets(np.array([301], np.uint64), cs, us)


def test_known_answer_two_records():
    """Hand-derived from the tabix layout: one reference, records at POS 100 and 20000, one 61-byte text block."""
    names = ["1"]
    vs = np.array([(500 << 16) | 10, (500 << 16) | 40], np.uint64)
    ve = np.array([(500 << 16) | 40, (561 << 16) | 0], np.uint64)
    got = tabix.build_tbi(names, np.array([0, 0]), np.array([100, 20000]), vs, ve)
    want = b"TBI\x01" + struct.pack("<8i", 1, 2, 1, 2, 0, 35, 0, 2) + b"1\0"
    want += struct.pack("<i", 3)
    want += struct.pack("<IiQQ", 4681, 1, int(vs[0]), int(ve[0]))          # (100-1) >> 14 = 0
    want += struct.pack("<IiQQ", 4682, 1, int(vs[1]), int(ve[1]))          # (20000-1) >> 14 = 1
    want += struct.pack("<IiQQQQ", 37450, 2, int(vs[0]), int(ve[1]), 2, 0)
    want += struct.pack("<i2Q", 2, int(vs[0]), int(vs[1]))
    want += struct.pack("<Q", 0)
    assert got == want
    assert tbi_reader.reg2bin(99, 100) == 4681 and tbi_reader.reg2bin(19999, 20000) == 4682


@pytest.mark.parametrize("seed,exact", [(3, False), (4, True), (5, False)])
def test_region_queries_match_brute_force(seed, exact):
    header, hb, body, bb, rows, labels, row_off = make_case(seed, exact_boundaries=exact)
    payload = build(hb, bb, rows, labels, row_off)
    tbi = tbi_reader.parse_tbi(payload)
    data = hb + bb + EOF_BLOCK
    assert tbi_reader.bgzf_inflate_all(data) == header + body
    assert (tbi["format"], tbi["col_seq"], tbi["col_beg"], tbi["col_end"], tbi["meta"], tbi["skip"]) == (2, 1, 2, 0, 35, 0)
    present = []
    for c, _ in rows:
        if c not in present:
            present.append(c)
    assert tbi["names"] == present and tbi["n_no_coor"] == 0
    for name, ref in zip(tbi["names"], tbi["refs"]):
        n = sum(1 for c, _ in rows if c == name)
        (first, last), (mapped, unmapped) = ref["bins"][tbi_reader.META_BIN]
        assert (mapped, unmapped) == (n, 0) and first < last
        assert ref["ioff"] == sorted(ref["ioff"]) and ref["ioff"][0] == first
        for b, chunks in ref["bins"].items():
            if b != tbi_reader.META_BIN:
                assert 4681 <= b < 37449 and len(chunks) == 1 and first <= chunks[0][0] < chunks[0][1] <= last
    rng = random.Random(seed + 100)
    for _ in range(60):
        chrom = rng.choice(labels + ["7"])
        beg = rng.randrange(0, 3_000_000)
        end = beg + rng.choice([1, 10, 5000, 20000, 400_000, 3_000_000])
        assert tbi_reader.query(data, tbi, chrom, beg, end) == brute(body, chrom, beg, end), (chrom, beg, end)
    # every single record can be found through the index
    for c, p in rows[::17]:
        assert brute(body, c, max(p - 1, 0), max(p, 1)) == tbi_reader.query(data, tbi, c, max(p - 1, 0), max(p, 1))


def test_position_zero_and_empty():
    t = tabix.build_tbi(["1"], np.array([0, 0]), np.array([0, 1]), np.array([5, 9], np.uint64), np.array([9, 20], np.uint64))
    ref = tbi_reader.parse_tbi(t)["refs"][0]
    assert ref["bins"][4681] == [(5, 20)] and ref["ioff"] == [5]
    empty = tbi_reader.parse_tbi(tabix.TabixBuilder().payload())
    assert empty["names"] == [] and empty["refs"] == []


def test_unindexable_orders_raise():
    v = np.arange(3, dtype=np.uint64)
    with pytest.raises(ValueError, match="not sorted"):
        tabix.build_tbi(["1"], np.array([0, 0, 0]), np.array([5, 4, 6]), v, v + 1)
    with pytest.raises(ValueError, match="not contiguous"):
        tabix.build_tbi(["1", "2"], np.array([0, 1, 0]), np.array([5, 6, 7]), v, v + 1)
    with pytest.raises(ValueError, match="address"):
        tabix.build_tbi(["1"], np.array([0]), np.array([(1 << 29) + 1]), v[:1], v[:1] + 1)
