"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

Takes reference-shaped objects (anything with the attributes of the reference's SNPTuples /
SampleInfo: ``id, chromosome, position, tuples`` and ``sex, is_control, deleterious_snps,
person_id``) so that tests read like the reference's own code, flattens them the way the
reference's row loop consumes them, and calls oracle/dnaf_oracle.c.

Reference lines restated here (host-side string work only):
  * row prefix  "%s\\t%i\\trs%s\\t%s\\t%s\\t40\\tPASS\\t.\\tGT\\t"     pop_factory.py:503-507
  * REF / ALT columns (ref_allele_tuple, alt_alleles)                  pop_factory.py:104-116
  * `snp.id not in sample.deleterious_snps` (dict membership, quirk R8) pop_factory.py:485
  * VCF header (gen_vcf_header)                                        pop_factory.py:36-44
  * .fam line (SampleInfo.to_fam_format)                               pop_factory.py:62-68
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libdnaf_oracle.so")
KMAX = 4
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "dnaf_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libdnaf_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        u8p = ctypes.POINTER(ctypes.c_uint8)
        u32p = ctypes.POINTER(ctypes.c_uint32)
        u64p = ctypes.POINTER(ctypes.c_uint64)
        f64p = ctypes.POINTER(ctypes.c_double)
        L.dnaf_or_philox4x32_10.argtypes = [u32p, u32p, u32p]
        L.dnaf_or_philox4x32_10.restype = None
        L.dnaf_or_uniform_bits.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32, u32p]
        L.dnaf_or_uniform_bits.restype = None
        L.dnaf_or_uniforms.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32, f64p]
        L.dnaf_or_uniforms.restype = None
        L.dnaf_or_pick_allele_index.argtypes = [f64p, ctypes.c_int, ctypes.c_double]
        L.dnaf_or_pick_allele_index.restype = ctypes.c_int
        L.dnaf_or_rows.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, u8p, u64p, u8p, u64p, u8p,
                                   f64p, ctypes.c_uint32, u8p, u8p, ctypes.c_uint64, u64p, u32p, u8p,
                                   ctypes.c_uint64, u64p, ctypes.c_int]
        L.dnaf_or_rows.restype = ctypes.c_int64
        L.dnaf_or_bgzf_bound.argtypes = [ctypes.c_uint64]
        L.dnaf_or_bgzf_bound.restype = ctypes.c_uint64
        L.dnaf_or_bgzf.argtypes = [u8p, ctypes.c_uint64, ctypes.c_int, ctypes.c_int, u8p, ctypes.c_uint64,
                                   ctypes.c_int]
        L.dnaf_or_bgzf.restype = ctypes.c_int64
        L.dnaf_or_crc32.argtypes = [u8p, ctypes.c_uint64]
        L.dnaf_or_crc32.restype = ctypes.c_uint32
        L.dnaf_or_num_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(ctypes.POINTER(ct))


def philox4x32_10(ctr, key):
    c = (ctypes.c_uint32 * 4)(*ctr)
    k = (ctypes.c_uint32 * 2)(*key)
    o = (ctypes.c_uint32 * 4)()
    lib().dnaf_or_philox4x32_10(c, k, o)
    return [int(x) for x in o]


def uniform_bits(seed, row, n_alleles):
    out = np.zeros(max(n_alleles, 1), dtype=np.uint32)
    lib().dnaf_or_uniform_bits(seed, row, n_alleles, _p(out, ctypes.c_uint32))
    return out[:n_alleles]


def uniforms(seed, row, n_alleles):
    out = np.zeros(max(n_alleles, 1), dtype=np.float64)
    lib().dnaf_or_uniforms(seed, row, n_alleles, _p(out, ctypes.c_double))
    return out[:n_alleles]


def pick_allele_index(cum, roll):
    c = np.asarray(cum, dtype=np.float64)
    r = lib().dnaf_or_pick_allele_index(_p(c, ctypes.c_double), len(c), float(roll))
    return None if r < 0 else r


# ----------------------------------------------------------------------------- host-side string work
def alt_alleles(snp):
    """pop_factory.py:111-116"""
    t = snp.tuples
    if len(t) == 1:
        return t[0][0]
    if len(t) == 2:
        return t[1][0]
    return ",".join(x[0] for x in t[1:])


def row_prefix(snp):
    """pop_factory.py:503-507"""
    return "%s\t%i\trs%s\t%s\t%s\t40\tPASS\t.\tGT\t" % (snp.chromosome, snp.position, snp.id, snp.tuples[0][0],
                                                          alt_alleles(snp))


def vcf_header(fam_data, filedate):
    """pop_factory.py:36-44 with the nondeterministic datetime.now() string passed in."""
    header = "##fileformat=VCFv4.3\n"
    header += "##filedate=%s\n" % filedate
    header += "##source=PopFactory\n"
    header += '##FILTER=<ID=q10,Description="Quality below 10">\n'
    header += '##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">\n'
    header += "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t"
    header += "\t".join(str(s.person_id) for s in fam_data) + "\n"
    return header


def fam_line(s):
    """pop_factory.py:62-68"""
    return "%i\t%i\t%i\t%i\t%i\t%i\t\n" % (s.family_id, s.person_id, s.father_id, s.mother_id, s.sex,
                                           1 if s.is_control else 2)


def override_pairs(fam_data, snps):
    """(local row, sample) pairs where the reference takes the forced-minor branch (:485, :495-499).

    Uses the reference's own membership test -- `snp.id in sample.deleterious_snps` on the dict as
    loaded -- so int-vs-str key behaviour is reproduced, not "fixed".
    """
    cases = [(i, s.deleterious_snps) for i, s in enumerate(fam_data) if not s.is_control]
    rows, samples = [], []
    if cases:
        keys = set()
        for _, d in cases:
            keys.update(d.keys())
        for r, snp in enumerate(snps):
            if snp.id in keys:
                for i, d in cases:
                    if snp.id in d:
                        rows.append(r)
                        samples.append(i)
    return np.asarray(rows, dtype=np.uint64), np.asarray(samples, dtype=np.uint32)


def flatten(fam_data, snps):
    """Flat arrays the C oracle consumes (also handy for tests of the CUDA boundary)."""
    n = len(fam_data)
    sex = np.asarray([s.sex for s in fam_data], dtype=np.uint8).reshape(n)
    ctl = np.asarray([1 if s.is_control else 0 for s in fam_data], dtype=np.uint8).reshape(n)
    chrom = [str(s.chromosome).encode("latin-1") for s in snps]
    prefix = [row_prefix(s).encode("latin-1") for s in snps]
    kk = np.asarray([len(s.tuples) for s in snps], dtype=np.uint8)
    if len(snps) and kk.max() > KMAX:
        raise ValueError("oracle supports at most %d alleles per SNP" % KMAX)
    cum = np.full((max(len(snps), 1), KMAX), 2.0, dtype=np.float64)
    for r, s in enumerate(snps):
        for k, t in enumerate(s.tuples):
            cum[r, k] = t[1]

    def cat(parts):
        off = np.zeros(len(parts) + 1, dtype=np.uint64)
        if parts:
            off[1:] = np.cumsum([len(x) for x in parts])
        data = np.frombuffer(b"".join(parts) + b"\0", dtype=np.uint8).copy()
        return data, off

    chrom_b, chrom_o = cat(chrom)
    prefix_b, prefix_o = cat(prefix)
    orow, osamp = override_pairs(fam_data, snps)
    return dict(n=n, sex=sex, is_control=ctl, chrom_bytes=chrom_b, chrom_off=chrom_o, prefix_bytes=prefix_b,
                prefix_off=prefix_o, K=kk, cum=cum, over_row=orow, over_sample=osamp)


def rows_from_flat(flat, seed, row_begin=0, n_threads=1):
    n_rows = len(flat["K"])
    n = flat["n"]
    cap = int(flat["prefix_off"][-1]) + n_rows * (4 * n + 1) + 16
    out = np.zeros(cap, dtype=np.uint8)
    row_off = np.zeros(n_rows + 1, dtype=np.uint64)
    sex = flat["sex"] if n else np.zeros(1, np.uint8)
    ctl = flat["is_control"] if n else np.zeros(1, np.uint8)
    orow = flat["over_row"] if len(flat["over_row"]) else np.zeros(1, np.uint64)
    osamp = flat["over_sample"] if len(flat["over_sample"]) else np.zeros(1, np.uint32)
    tot = lib().dnaf_or_rows(seed, row_begin, n_rows, _p(flat["chrom_bytes"], ctypes.c_uint8),
                             _p(flat["chrom_off"], ctypes.c_uint64), _p(flat["prefix_bytes"], ctypes.c_uint8),
                             _p(flat["prefix_off"], ctypes.c_uint64), _p(flat["K"], ctypes.c_uint8),
                             _p(flat["cum"], ctypes.c_double), n, _p(sex, ctypes.c_uint8),
                             _p(ctl, ctypes.c_uint8), len(flat["over_row"]), _p(orow, ctypes.c_uint64),
                             _p(osamp, ctypes.c_uint32), _p(out, ctypes.c_uint8), cap,
                             _p(row_off, ctypes.c_uint64), n_threads)
    if tot == -1:
        raise TypeError("%i format: a real number is required, not NoneType")  # what the reference raises
    if tot < 0:
        raise RuntimeError("oracle buffer too small")
    return out[:tot], row_off


def rows(fam_data, snps, seed, row_begin=0, n_threads=1):
    """Text of VCF data rows for `snps` (global row index of snps[0] = row_begin) -> (bytes, row_off)."""
    text, row_off = rows_from_flat(flatten(fam_data, snps), seed, row_begin, n_threads)
    return text.tobytes(), row_off


def bgzf(text, level=6, with_eof=True, n_threads=1):
    buf = np.frombuffer(bytes(text) + b"\0", dtype=np.uint8)
    n = len(buf) - 1
    cap = int(lib().dnaf_or_bgzf_bound(n))
    out = np.zeros(cap, dtype=np.uint8)
    got = lib().dnaf_or_bgzf(_p(buf, ctypes.c_uint8), n, level, 1 if with_eof else 0, _p(out, ctypes.c_uint8), cap,
                             n_threads)
    if got < 0:
        raise RuntimeError("oracle bgzf failed")
    return out[:got].tobytes()


def bgzf_decompress(data):
    """Walk a BGZF stream block by block, checking framing, CRC32 and ISIZE -> (text, n_blocks, saw_eof)."""
    import struct
    import zlib
    out = []
    pos = 0
    blocks = 0
    last_isize = None
    data = bytes(data)
    while pos < len(data):
        if data[pos:pos + 4] != b"\x1f\x8b\x08\x04":
            raise ValueError("bad BGZF magic at %d" % pos)
        xlen = struct.unpack_from("<H", data, pos + 10)[0]
        if xlen != 6 or data[pos + 12:pos + 16] != b"BC\x02\x00":
            raise ValueError("bad BGZF extra field at %d" % pos)
        bsize = struct.unpack_from("<H", data, pos + 16)[0] + 1
        if pos + bsize > len(data):
            raise ValueError("truncated BGZF block at %d" % pos)
        cdata = data[pos + 18:pos + bsize - 8]
        crc, isize = struct.unpack_from("<II", data, pos + bsize - 8)
        d = zlib.decompressobj(-15)
        raw = d.decompress(cdata) + d.flush()
        if not d.eof or d.unused_data:
            raise ValueError("deflate stream of block at %d does not end cleanly" % pos)
        if len(raw) != isize:
            raise ValueError("ISIZE mismatch in block at %d: %d != %d" % (pos, len(raw), isize))
        if (zlib.crc32(raw) & 0xFFFFFFFF) != crc:
            raise ValueError("CRC mismatch in block at %d" % pos)
        if isize > 65536:
            raise ValueError("block at %d holds more than 64 KiB" % pos)
        out.append(raw)
        last_isize = isize
        blocks += 1
        pos += bsize
    return b"".join(out), blocks, last_isize == 0
