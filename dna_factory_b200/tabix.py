"""Tabix (.tbi) index of population.vcf.gz, built while the file is written.

The reference leaves indexing to a later `bcftools index` pass over the finished file (README.md:98-99), which has to
inflate all of it again.  The writer here already knows everything an index holds -- the length of every row
(`dnaf_row_offsets`, i.e. what queue_vcf_snps pop_factory.py:503-508 would have produced) and the compressed / text
size of every BGZF block it appended (`dnaf_block_log`, `dnaf_bgzf_scan`) -- so the index is a few numpy passes
(SURVEY 8f-3).  Layout follows the tabix specification (TBI\\1, VCF preset: sequence column 1, begin column 2,
meta '#'; UCSC binning with 16 KiB leaves; 16 KiB linear index; htslib's pseudo-bin 37450 with the record counts).

A SNP row covers one base (REF is one nucleotide), [POS-1, POS) 0-based, so it always falls into a leaf bin
4681 + ((POS-1) >> 14) and, rows being sorted by position inside a chromosome, every bin is one chunk.
`tests/tbi_reader.py` is an independent reader of the format used to check region queries against a brute-force
scan of the inflated text.  Parity note: no tabix / htslib binary exists in this image, so the index is checked
against the specification as restated in that reader, not against htslib output.
"""
import struct

import numpy as np

TBX_VCF = 2
LEAF_BIN0 = 4681          # ((1 << 15) - 1) // 7: first bin of the 16 KiB level
LEAF_SHIFT = 14
META_BIN = 37450          # htslib: one past the last real bin; holds (first, last offset) and (n records, 0)
MAX_POS = 1 << 29         # .tbi cannot address beyond 512 Mbp


def virtual_offsets(text_off, blk_csize, blk_usize):
    """BGZF virtual offsets (block start in the file << 16 | offset inside the inflated block) of absolute offsets
    `text_off` into the inflated stream.  An offset that falls on a block boundary belongs to the block that starts
    there (what bgzf_tell reports after reading up to the boundary); the end of the stream maps to the start of
    whatever follows the listed blocks (the EOF block)."""
    text_off = np.asarray(text_off, dtype=np.uint64)
    ustart = np.zeros(len(blk_usize) + 1, np.uint64)
    np.cumsum(blk_usize, dtype=np.uint64, out=ustart[1:])
    cstart = np.zeros(len(blk_csize) + 1, np.uint64)
    np.cumsum(blk_csize, dtype=np.uint64, out=cstart[1:])
    if len(text_off) and int(text_off.max()) > int(ustart[-1]):
        raise ValueError("offset beyond the end of the block table")
    b = np.searchsorted(ustart, text_off, side="right") - 1
    b = np.minimum(b, len(blk_usize))          # the end of the stream: one past the last block
    within = text_off - ustart[b]
    if len(within) and int(within.max()) >= 1 << 16:
        raise ValueError("block table and offsets disagree")
    return (cstart[b] << np.uint64(16)) | within


class TabixBuilder:
    """Collects the block table and the row table as the writer goes; payload() is the uncompressed .tbi."""

    def __init__(self):
        self._cs, self._us = [], []
        self._text = 0
        self._rows = []

    @property
    def text_bytes(self):
        return self._text

    def add_blocks(self, csize, usize):
        """Blocks appended to the file, in file order."""
        csize, usize = np.asarray(csize, np.uint32), np.asarray(usize, np.uint32)
        if len(csize) != len(usize):
            raise ValueError("block table columns differ in length")
        self._cs.append(csize)
        self._us.append(usize)
        self._text += int(usize.sum(dtype=np.uint64))

    def add_rows(self, chrom_labels, chrom_idx, position, row_off):
        """Rows about to be appended: `row_off[i]` = text offset of row i from the CURRENT end of the text (length
        S + 1, the last entry is the end of the last row).  Call before add_blocks() of the blocks that hold them."""
        row_off = np.asarray(row_off, np.uint64)
        if len(row_off) != len(position) + 1:
            raise ValueError("row_off must have one entry per row plus the end")
        self._rows.append((list(chrom_labels), np.asarray(chrom_idx, np.int64), np.asarray(position, np.int64),
                           row_off + np.uint64(self._text)))

    def payload(self):
        cs = np.concatenate(self._cs) if self._cs else np.zeros(0, np.uint32)
        us = np.concatenate(self._us) if self._us else np.zeros(0, np.uint32)
        names, name_id = [], {}
        chrom, pos, start, end = [], [], [], []
        for labels, ci, p, off in self._rows:
            if not len(p):
                continue
            remap = np.empty(len(labels), np.int64)
            # names in order of first appearance in the file
            first = np.unique(ci, return_index=True)
            for k in first[0][np.argsort(first[1])]:
                lab = str(labels[int(k)])
                if lab not in name_id:
                    name_id[lab] = len(names)
                    names.append(lab)
                remap[int(k)] = name_id[lab]
            chrom.append(remap[ci])
            pos.append(p)
            start.append(off[:-1])
            end.append(off[1:])
        if chrom:
            chrom, pos = np.concatenate(chrom), np.concatenate(pos)
            start, end = np.concatenate(start), np.concatenate(end)
        else:
            chrom, pos = np.zeros(0, np.int64), np.zeros(0, np.int64)
            start = end = np.zeros(0, np.uint64)
        return build_tbi(names, chrom, pos, virtual_offsets(start, cs, us), virtual_offsets(end, cs, us))


def validate_rows(chrom_idx, position):
    """What a .tbi index needs of the rows, checked BEFORE the (long) generation instead of after it: positions
    within 0 .. 2^29, every chromosome's rows contiguous, positions sorted inside a chromosome.  Raises ValueError."""
    chrom = np.asarray(chrom_idx, np.int64)
    position = np.asarray(position, np.int64)
    S = len(chrom)
    if not S:
        return
    if int(position.min()) < 0 or int(position.max()) > MAX_POS:
        raise ValueError("position outside what a .tbi index can address (0 .. 2^29)")
    cut = np.flatnonzero(np.diff(chrom)) + 1
    seen = chrom[np.concatenate(([0], cut))]
    if len(np.unique(seen)) != len(seen):
        raise ValueError("rows of one chromosome are not contiguous: the file cannot be tabix-indexed")
    same = np.diff(chrom) == 0
    if np.any(np.diff(np.maximum(position - 1, 0))[same] < 0):
        raise ValueError("rows of a chromosome are not sorted by position: the file cannot be tabix-indexed")


def build_tbi(names, chrom, position, voff_start, voff_end):
    """Uncompressed .tbi bytes.  `chrom[r]` indexes `names`; rows must be grouped by chromosome and sorted by
    position inside each (what tabix requires of the file; pop_factory.py:245 sorts that way)."""
    chrom = np.asarray(chrom, np.int64)
    position = np.asarray(position, np.int64)
    S = len(chrom)
    if S and (int(position.min()) < 0 or int(position.max()) > MAX_POS):
        raise ValueError("position outside what a .tbi index can address (0 .. 2^29)")
    cut = np.flatnonzero(np.diff(chrom)) + 1 if S else np.zeros(0, np.int64)
    bounds = np.concatenate(([0], cut, [S])).astype(np.int64) if S else np.zeros(1, np.int64)
    seen = chrom[bounds[:-1]] if S else np.zeros(0, np.int64)
    if len(np.unique(seen)) != len(seen):
        raise ValueError("rows of one chromosome are not contiguous: the file cannot be tabix-indexed")
    names_blob = b"".join(n.encode("latin-1") + b"\0" for n in names)
    out = [b"TBI\x01", struct.pack("<8i", len(names), TBX_VCF, 1, 2, 0, ord("#"), 0, len(names_blob)), names_blob]
    per_ref = {int(c): (int(a), int(b)) for c, a, b in zip(seen, bounds[:-1], bounds[1:])}
    bin_t = np.dtype([("bin", "<u4"), ("n_chunk", "<i4"), ("beg", "<u8"), ("end", "<u8")])
    for ref in range(len(names)):
        if ref not in per_ref:
            out.append(struct.pack("<ii", 0, 0))
            continue
        a, b = per_ref[ref]
        beg0 = np.maximum(position[a:b] - 1, 0)          # htslib clamps POS 0 to the first base
        if np.any(np.diff(beg0) < 0):
            raise ValueError("rows of chromosome %s are not sorted by position" % names[ref])
        win = beg0 >> LEAF_SHIFT
        first = np.concatenate(([0], np.flatnonzero(np.diff(win)) + 1))       # first row of every occupied window
        last = np.concatenate((first[1:], [b - a])) - 1                       # its last row
        bins = np.empty(len(first), bin_t)
        bins["bin"] = LEAF_BIN0 + win[first]
        bins["n_chunk"] = 1
        bins["beg"] = voff_start[a:b][first]
        bins["end"] = voff_end[a:b][last]
        out.append(struct.pack("<i", len(first) + 1))
        out.append(bins.tobytes())
        out.append(struct.pack("<IiQQQQ", META_BIN, 2, int(voff_start[a]), int(voff_end[b - 1]), b - a, 0))
        n_intv = int(win[-1]) + 1
        ioff = np.full(n_intv, np.iinfo(np.uint64).max, np.uint64)
        ioff[win[first]] = voff_start[a:b][first]
        ioff = np.minimum.accumulate(ioff[::-1])[::-1]                        # empty windows point at the next record
        out.append(struct.pack("<i", n_intv))
        out.append(ioff.astype("<u8").tobytes())
    out.append(struct.pack("<Q", 0))                                          # records without coordinates
    return b"".join(out)
