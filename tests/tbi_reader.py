"""Independent reader of the tabix (.tbi) format + BGZF random access, for the index tests.

Written from the tabix specification (header, per-reference bins / chunks / 16 KiB linear index, UCSC binning
scheme reg2bins) -- it shares no code with dna_factory_b200/tabix.py.  query() answers a region the way a tabix
client does: bins overlapping the region -> their chunks -> drop chunks that end before the linear-index offset
of the region's first window -> seek to each chunk's virtual offset and read lines until the chunk ends.
"""
import struct
import zlib

META_BIN = 37450


def bgzf_inflate_all(data):
    """Whole BGZF stream -> bytes (block by block, checking ISIZE and CRC32)."""
    out, o = [], 0
    while o < len(data):
        text, csize = read_block(data, o)
        out.append(text)
        o += csize
    return b"".join(out)


def read_block(data, coffset):
    """(inflated bytes, compressed size) of the BGZF block that starts at file offset `coffset`."""
    h = data[coffset:coffset + 18]
    assert h[:4] == b"\x1f\x8b\x08\x04" and h[12:14] == b"BC", "not a BGZF block at %d" % coffset
    csize = struct.unpack_from("<H", h, 16)[0] + 1
    body = data[coffset + 18:coffset + csize - 8]
    crc, isize = struct.unpack_from("<II", data, coffset + csize - 8)
    text = zlib.decompress(body, -15)
    assert len(text) == isize and zlib.crc32(text) == crc
    return text, csize


def parse_tbi(payload):
    magic, n_ref, fmt, col_seq, col_beg, col_end, meta, skip, l_nm = struct.unpack_from("<4s8i", payload, 0)
    assert magic == b"TBI\x01"
    o = 36
    names = payload[o:o + l_nm].split(b"\0")[:-1]
    assert len(names) == n_ref
    o += l_nm
    refs = []
    for _ in range(n_ref):
        n_bin = struct.unpack_from("<i", payload, o)[0]
        o += 4
        bins = {}
        for _ in range(n_bin):
            b, n_chunk = struct.unpack_from("<Ii", payload, o)
            o += 8
            chunks = [struct.unpack_from("<QQ", payload, o + 16 * k) for k in range(n_chunk)]
            o += 16 * n_chunk
            assert b not in bins
            bins[b] = chunks
        n_intv = struct.unpack_from("<i", payload, o)[0]
        o += 4
        ioff = list(struct.unpack_from("<%dQ" % n_intv, payload, o))
        o += 8 * n_intv
        refs.append({"bins": bins, "ioff": ioff})
    n_no_coor = None
    if o + 8 <= len(payload):
        n_no_coor = struct.unpack_from("<Q", payload, o)[0]
        o += 8
    assert o == len(payload), "trailing bytes in the index"
    return {"format": fmt, "col_seq": col_seq, "col_beg": col_beg, "col_end": col_end, "meta": meta, "skip": skip,
            "names": [n.decode() for n in names], "refs": refs, "n_no_coor": n_no_coor}


def reg2bins(beg, end):
    """Bins that may hold records overlapping 0-based half-open [beg, end)."""
    end -= 1
    bins = [0]
    for shift, first in ((26, 1), (23, 9), (20, 73), (17, 585), (14, 4681)):
        bins.extend(range(first + (beg >> shift), first + (end >> shift) + 1))
    return bins


def reg2bin(beg, end):
    end -= 1
    for shift, first in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        if beg >> shift == end >> shift:
            return first + (beg >> shift)
    return 0


class BgzfCursor:
    """Sequential line reader positioned by virtual offset."""

    def __init__(self, data, voff):
        self.data = data
        self.coffset, self.within = voff >> 16, voff & 0xFFFF
        self.text, self.csize = read_block(data, self.coffset) if self.coffset < len(data) else (b"", 0)

    def tell(self):
        # what bgzf_tell reports: once a block is used up the position is the start of the next one
        self._roll()
        return (self.coffset << 16) | self.within

    def _roll(self):
        while self.csize and self.within >= len(self.text):
            self.coffset += self.csize
            self.within = 0
            self.text, self.csize = read_block(self.data, self.coffset) if self.coffset < len(self.data) else (b"", 0)

    def readline(self):
        parts = []
        while True:
            self._roll()
            if not self.csize:
                return b"".join(parts)
            nl = self.text.find(b"\n", self.within)
            if nl >= 0:
                parts.append(self.text[self.within:nl + 1])
                self.within = nl + 1
                return b"".join(parts)
            parts.append(self.text[self.within:])
            self.within = len(self.text)


def query(data, tbi, chrom, beg, end):
    """Lines of the BGZF VCF `data` on `chrom` whose base [POS-1, POS) overlaps 0-based half-open [beg, end)."""
    if chrom not in tbi["names"]:
        return []
    ref = tbi["refs"][tbi["names"].index(chrom)]
    w = beg >> 14
    min_off = ref["ioff"][w] if w < len(ref["ioff"]) else (ref["ioff"][-1] if ref["ioff"] else 0)
    chunks = []
    for b in reg2bins(beg, end):
        if b in ref["bins"] and b != META_BIN:
            chunks.extend(c for c in ref["bins"][b] if c[1] > min_off)
    chunks.sort()
    hits = []
    name = chrom.encode()
    for c_beg, c_end in chunks:
        cur = BgzfCursor(data, c_beg)
        while cur.tell() < c_end:
            line = cur.readline()
            if not line:
                break
            f = line.split(b"\t", 2)
            assert f[0] == name, "chunk of %s holds a %r record" % (chrom, f[0])
            p0 = max(int(f[1]) - 1, 0)
            if p0 >= end:
                break
            if p0 + 1 > beg:
                hits.append(line)
    return hits
