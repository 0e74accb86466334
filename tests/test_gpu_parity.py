"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path, called through the C ABI,
against the committed reference goldens and against the CPU oracle on seeded synthetic inputs."""
import os
import zlib

import numpy as np
import pytest

from tests.cases import ROW_CASES, load_case, synth_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def native():
    from dna_factory_b200 import _native, host
    return _native, host


def _engine(native, case, chunk=None, fused=True):
    _native, host = native
    eng = _native.Engine(0)
    host.configure(eng, case.samples, case.snps)
    if chunk:
        eng.set_chunk_bytes(chunk)
    eng.set_fused(fused)
    return eng


def _expected_genotypes(case, text):
    """Parse the oracle/golden text back into the [rows][N][2] allele matrix (0xFF = absent)."""
    n = len(case.samples)
    rows = text.split(b"\n")[:-1]
    out = np.full((len(rows), n, 2), 0xFF, dtype=np.uint8)
    for r, line in enumerate(rows):
        cells = line.split(b"\t")[9:]
        if n == 0:
            continue
        assert len(cells) == n
        for i, c in enumerate(cells):
            if c == b".":
                continue
            parts = c.split(b"/")
            out[r, i, 0] = int(parts[0])
            if len(parts) == 2:
                out[r, i, 1] = int(parts[1])
    return out


@pytest.mark.parametrize("name", ROW_CASES)
def test_genotypes_bit_exact_vs_reference_golden(native, name):
    case = load_case(name)
    if len(case.samples) == 0:
        pytest.skip("no genotypes in a sites-only run")
    eng = _engine(native, case)
    eng.set_row_base(case.row_begin)
    got = eng.genotypes(0, len(case.snps), case.seed)
    assert np.array_equal(got, _expected_genotypes(case, case.text))


@pytest.mark.parametrize("name", ROW_CASES)
def test_text_byte_identical_vs_reference_golden(native, name):
    case = load_case(name)
    eng = _engine(native, case)
    eng.set_row_base(case.row_begin)
    assert eng.text(0, len(case.snps), case.seed) == case.text


@pytest.mark.parametrize("name", ROW_CASES)
@pytest.mark.parametrize("level", [1, 6, 9])
def test_bgzf_decompresses_to_reference_golden(native, name, level):
    import gzip
    from oracle import oracle
    case = load_case(name)
    eng = _engine(native, case)
    eng.set_row_base(case.row_begin)
    blob, st = eng.generate(0, len(case.snps), case.seed, level=level)
    text, blocks, eof = oracle.bgzf_decompress(blob)
    assert text == case.text
    assert blocks == st["bgzf_blocks"] and not eof
    assert st["text_bytes"] == len(case.text)
    assert gzip.decompress(blob + native[0].bgzf_eof()) == case.text


@pytest.mark.parametrize("n,s,seed", [(1, 5, 1), (15, 9, 2), (16, 9, 3), (17, 40, 4), (200, 300, 5), (1000, 64, 6),
                                      (4097, 24, 7), (20000, 6, 8)])
def test_oracle_vs_cuda_synthetic(native, n, s, seed):
    from oracle import oracle
    case = synth_case(n, s, seed=seed)
    want, row_off = oracle.rows(case.samples, case.snps, case.seed, 0, n_threads=4)
    eng = _engine(native, case, chunk=1 << 20)
    assert np.array_equal(eng.genotypes(0, s, case.seed), _expected_genotypes(case, want))
    assert eng.text(0, s, case.seed) == want
    for level in (1, 6):
        blob, st = eng.generate(0, s, case.seed, level=level)
        text, blocks, _ = oracle.bgzf_decompress(blob)
        assert text == want
        assert st["calls"] == n * s and st["bgzf_blocks"] == blocks
    # any row range is independently regenerable (counter-based RNG): ranges concatenate
    mid = s // 2
    a, _ = eng.generate(0, mid, case.seed, level=2)
    b, _ = eng.generate(mid, s, case.seed, level=2)
    assert oracle.bgzf_decompress(a + b)[0] == want
    sub = eng.text(mid, s, case.seed)
    assert sub == want[int(row_off[mid]):]


@pytest.mark.parametrize("n,s,seed", [(4096, 12, 31), (4097, 12, 32), (4160, 8, 33), (16256, 5, 34), (16257, 5, 35),
                                      (16320, 4, 36), (20000, 24, 37), (33000, 6, 38), (100003, 5, 39)])
def test_fused_kernel_vs_oracle_and_generic(native, n, s, seed):
    """The fused sample+format+deflate kernel (autosome rows of >= 4096 samples) against the oracle text and
    against the three-kernel path; block streams may differ, decompressed bytes may not."""
    from oracle import oracle
    case = synth_case(n, s, seed=seed, chroms=['1', '2', '1', '7', 'X', '22', 'Y'], n_del=40)
    want, _ = oracle.rows(case.samples, case.snps, case.seed, 0, n_threads=4)
    eng = _engine(native, case, chunk=8 << 20)
    fused, st_f = eng.generate(0, s, case.seed, level=2)
    assert st_f["ms_fused"] > 0
    assert oracle.bgzf_decompress(fused)[0] == want
    eng.set_fused(False)
    plain, st_p = eng.generate(0, s, case.seed, level=2)
    assert st_p["ms_fused"] == 0
    assert oracle.bgzf_decompress(plain)[0] == want
    # the static per-MAF-bucket codes must stay close to the per-block dynamic codes
    assert st_f["bgzf_bytes"] < 1.10 * st_p["bgzf_bytes"] + 4096


@pytest.mark.parametrize("n,s,seed", [(4096, 10, 51), (4160, 8, 52), (16256, 5, 53), (16257, 5, 54), (20000, 24, 55),
                                      (33000, 6, 56), (100003, 4, 57)])
def test_lz_tiers_decompress_to_the_oracle_rows(native, n, s, seed):
    """-z 3..9 run the LZ77 kernel k_lz on autosome rows (pop_factory.py:403 hands -z to the writer): every level must
    decompress to the oracle's rows, be deterministic (two runs, same bytes), and a deeper tier must not be larger."""
    from oracle import oracle
    case = synth_case(n, s, seed=seed, chroms=['1', '2', '1', '7', 'X', '22', 'Y'], n_del=40)
    want, _ = oracle.rows(case.samples, case.snps, case.seed, 0, n_threads=4)
    eng = _engine(native, case, chunk=8 << 20)
    sizes = {}
    for level in (2, 3, 4, 5, 6, 7, 8, 9):
        blob, st = eng.generate(0, s, case.seed, level=level)
        text, blocks, _ = oracle.bgzf_decompress(blob)
        assert text == want, "level %d" % level
        assert blocks == st["bgzf_blocks"] and st["ms_fused"] > 0
        again, _ = eng.generate(0, s, case.seed, level=level)
        assert again == blob, "level %d is not deterministic" % level
        sizes[level] = len(blob)
    # a handful of rows: allow 1.5 % noise between neighbouring tiers (the large-sample check is bench.py's level_sweep)
    assert all(sizes[b] <= 1.015 * sizes[a] for a, b in zip((2, 3, 4, 5, 6, 7, 8), (3, 4, 5, 6, 7, 8, 9))), sizes
    assert sizes[9] <= sizes[4] <= sizes[2], sizes
    if s >= 20:
        assert sizes[6] < 0.85 * sizes[2], sizes


@pytest.mark.parametrize("level", [3, 4, 6, 9])
def test_lz_tiers_dense_overrides(native, level):
    """The dense forced-minor patterns of test_auto_kernel_dense_overrides_on_rare_rows through k_lz: spans that
    overflow their staging words re-emit straight into the block, every symbol has a code."""
    from oracle import oracle
    from tests.cases import Snp, Sample
    n, stride = 12345, 1
    snps = [Snp(id=i + 1, chromosome='1', position=1000 * (i + 1), tuples=[("A", 1 - maf), ("C", 1.0)])
            for i, maf in enumerate([0.005, 0.01, 0.02, 0.25, 0.495, 0.005])]
    samples = []
    for i in range(n):
        ctl = i < n // 3
        d = None
        if not ctl:
            d = {sn.id: 0.5 for sn in snps[:5]} if (i % stride == 0 and (i // 97) % 2 == 0) else {}
            if n // 2 <= i < n // 2 + 70:
                d[6] = 0.5
        samples.append(Sample(family_id=i + 1, person_id=100001 + i, father_id=0, mother_id=0, sex=1 + (i & 1),
                              is_control=ctl, deleterious_snps=d))
    from types import SimpleNamespace
    case = SimpleNamespace(name="dense", seed=0xD15EA5E, row_begin=0, samples=samples, snps=snps, text=None)
    want, _ = oracle.rows(case.samples, case.snps, case.seed, 0, n_threads=4)
    eng = _engine(native, case)
    blob, st = eng.generate(0, len(snps), case.seed, level=level)
    assert oracle.bgzf_decompress(blob)[0] == want


@pytest.mark.parametrize("n,stride", [(8192, 2), (8200, 3), (20000, 7), (12345, 1)])
def test_auto_kernel_dense_overrides_on_rare_rows(native, n, stride):
    """Forced-minor cells (pop_factory.py:495-499) in patterns the rare-MAF code tables never expect: byte patterns
    whose tokens do not fit a LUT entry and spans that overflow their staging words must still decode exactly."""
    from oracle import oracle
    from tests.cases import Snp, Sample
    from types import SimpleNamespace
    snps = [Snp(id=i + 1, chromosome='1', position=1000 * (i + 1), tuples=[("A", 1 - maf), ("C", 1.0)])
            for i, maf in enumerate([0.005, 0.01, 0.02, 0.25, 0.495, 0.005])]
    # X rows take k_x: the same forced patterns on a rare row, a common row and a single-allele row
    snps += [Snp(id=7, chromosome='X', position=5, tuples=[("A", 0.995), ("C", 1.0)]),
             Snp(id=8, chromosome='X', position=6, tuples=[("A", 0.6), ("C", 1.0)]),
             Snp(id=9, chromosome='X', position=7, tuples=[("G", 1.0)])]
    samples = []
    for i in range(n):
        ctl = i < n // 3
        d = None
        if not ctl:
            # every `stride`-th case carries every SNP; the last row only on a short burst of cases
            d = {s.id: 0.5 for s in snps[:5] + snps[6:]} if (i % stride == 0) else {}
            if n // 2 <= i < n // 2 + 70:
                d[6] = 0.5
        samples.append(Sample(family_id=i + 1, person_id=100001 + i, father_id=0, mother_id=0, sex=1 + (i & 1),
                              is_control=ctl, deleterious_snps=d))
    case = SimpleNamespace(name="dense", seed=0xD15EA5E, row_begin=0, samples=samples, snps=snps, text=None)
    want, _ = oracle.rows(case.samples, case.snps, case.seed, 0, n_threads=4)
    eng = _engine(native, case)
    blob, st = eng.generate(0, len(snps), case.seed, level=2)
    assert st["ms_fused"] > 0
    text, blocks, _ = oracle.bgzf_decompress(blob)
    assert text == want and blocks == st["bgzf_blocks"]


@pytest.mark.parametrize("n,s,seed,odds", [(8192, 16, 41, 0.5), (8193, 16, 42, 0.5), (9000, 12, 43, 0.0),
                                           (9000, 12, 44, 1.0), (20000, 20, 45, 0.5), (40000, 8, 46, 0.3)])
def test_fused_text_kernel_all_classes(native, n, s, seed, odds):
    """X / Y / MT rows and multi-allelic autosome rows go through k_fused_text (text staged in shared memory)."""
    from oracle import oracle
    case = synth_case(n, s, seed=seed, male_odds=odds, chroms=['X', 'Y', 'MT', '3', 'X'], n_del=60, exotic=True)
    want, _ = oracle.rows(case.samples, case.snps, case.seed, 0, n_threads=4)
    eng = _engine(native, case, chunk=16 << 20)
    fused, st_f = eng.generate(0, s, case.seed, level=2)
    assert st_f["ms_fused"] > 0 and st_f["ms_sample"] == 0
    assert oracle.bgzf_decompress(fused)[0] == want
    eng.set_fused(False)
    plain, st_p = eng.generate(0, s, case.seed, level=2)
    assert oracle.bgzf_decompress(plain)[0] == want
    # rows with 3-4 alleles share the biallelic MAF buckets, so the static codes fit them less well
    assert st_f["bgzf_bytes"] < 1.35 * st_p["bgzf_bytes"] + 8192


def test_bgzf_compress_arbitrary_bytes(native):
    from oracle import oracle
    _native, _ = native
    eng = _native.Engine(0)
    rs = np.random.RandomState(0)
    header = ("##fileformat=VCFv4.3\n#CHROM\tPOS\t" + "\t".join(str(100001 + i) for i in range(30000)) + "\n").encode()
    cases = [b"", b"a", b"abcd" * 5, header, rs.randint(0, 256, 200000).astype(np.uint8).tobytes(),
             b"0/0\t" * 70000, bytes(range(256)) * 300]
    for data in cases:
        blob, st = eng.bgzf_compress(data)
        text, blocks, _ = oracle.bgzf_decompress(blob)
        assert text == data
        assert blocks == (len(data) + 65279) // 65280


def test_stream_sink_and_chunking(native):
    from oracle import oracle
    case = synth_case(3000, 50, seed=11)
    want, _ = oracle.rows(case.samples, case.snps, case.seed, 0, n_threads=4)
    eng = _engine(native, case, chunk=64 << 10)
    pieces = []
    st = eng.generate_stream(0, 50, case.seed, pieces.append, level=3)
    assert len(pieces) > 1
    assert oracle.bgzf_decompress(b"".join(pieces))[0] == want
    dev = eng.generate_device(0, 50, case.seed, level=3)
    assert dev["bgzf_bytes"] == st["bgzf_bytes"] and dev["crc_xor"] == st["crc_xor"]
    assert dev["kernel_launches"] > 0
    # checksum of checksums: xor of the per-block CRC32s in the stream
    x = 0
    blob = b"".join(pieces)
    pos = 0
    while pos < len(blob):
        bsize = int.from_bytes(blob[pos + 16:pos + 18], "little") + 1
        x ^= int.from_bytes(blob[pos + bsize - 8:pos + bsize - 4], "little")
        pos += bsize
    assert x == st["crc_xor"]


def test_fd_sink_writes_the_same_stream(native, tmp_path):
    """dnaf_generate_fd (the library writes the blocks itself) against the buffer sink: identical bytes on disk."""
    case = synth_case(6000, 40, seed=12)
    eng = _engine(native, case, chunk=4 << 20)
    want, st0 = eng.generate(0, 40, case.seed, level=2)
    path = tmp_path / "rows.bgzf"
    with open(path, "wb") as f:
        f.write(b"HEAD")
        f.flush()
        st = eng.generate_fd(0, 40, case.seed, f.fileno(), level=2)
    assert path.read_bytes() == b"HEAD" + want
    assert st["bgzf_bytes"] == len(want) and st["crc_xor"] == st0["crc_xor"]
    with pytest.raises(native[0].DnafError):
        eng.generate_fd(0, 40, case.seed, 10 ** 6, level=2)      # not an open descriptor


def test_errors_are_loud(native):
    _native, host = native
    eng = _native.Engine(0)
    with pytest.raises(_native.DnafError):
        eng.generate(0, 1, 1)                      # nothing configured
    case = synth_case(4, 3, seed=1)
    host.configure(eng, case.samples, case.snps)
    with pytest.raises(_native.DnafError):
        eng.generate(0, 99, 1)                     # row range out of bounds
    flat = host.flatten_snps(case.snps)
    flat["thresholds"] = flat["thresholds"].copy()
    flat["thresholds"][0, 1] = 5                   # CDF does not reach 1.0 -> the reference would raise
    with pytest.raises(_native.DnafError) as e:
        eng.set_snps(**flat)
    assert e.value.code == _native.E_INPUT


@pytest.mark.parametrize("size,min_maf,seed", [(1, 0.005, 3), (1000, 0.01, 7), (100000, 0.16, 0x5EED000000000001),
                                               ((1 << 20) + 17, 0.01, 99)])
def test_snp_selection_device_vs_oracle(native, size, min_maf, seed):
    """dnaf_select_snps (inverse-CDF sampler + radix sort) against the numpy restatement of
    SnpFactory.random_snp_tuples / the sort of pop_factory.py:245: every column identical, sorted and unsorted."""
    from dna_factory_b200 import snp
    from oracle import snp_select
    _native, _ = native
    eng = _native.Engine(0)
    fac = snp.SnpFactory.init_from_cdf_file()
    t = fac.selection_tables(min_maf)
    for sort in (True, False):
        got = eng.select_snps(size, seed, t["chrom_cdf"], t["chrom_max_pos"], t["chrom_rank"], t["maf_cdf"], sort=sort)
        want = snp_select.select(seed, size, t["chrom_cdf"], t["chrom_max_pos"], t["chrom_rank"], t["maf_cdf"], sort=sort)
        for k in want:
            assert np.array_equal(got[k], want[k]), k
    # the reference's own tolerances (test/unit/snp_factory_test.py:14-37) on the device draw
    if size >= 100000:
        tab = fac.random_snp_table_device(eng, size, seed, min_maf=min_maf)
        assert np.all(1 - tab.cum[:, 0] >= min_maf - 1e-12) and not np.any(tab.nts[:, 0] == tab.nts[:, 1])
        assert abs(np.mean(tab.chrom_idx == 0) - snp.CHROMOSOME_PROB[0]) < 0.01


def test_snp_selection_device_feeds_generation(native):
    """Selection on the GPU -> the usual flat arrays -> rows: the decompressed VCF equals the oracle's rows for the
    same SNP list (the drop-in flow of a --generate_snps run)."""
    from dna_factory_b200 import snp
    from oracle import oracle
    _native, host = native
    eng = _native.Engine(0)
    fac = snp.SnpFactory.init_from_cdf_file()
    tab = fac.random_snp_table_device(eng, 300, 1234, min_maf=0.01)
    case = synth_case(5000, 3, seed=5)
    snps = tab.to_snps()
    want, _ = oracle.rows(case.samples, snps, 77, 0, n_threads=4)
    host.configure(eng, case.samples, snps)
    blob, st = eng.generate(0, len(snps), 77, level=2)
    assert oracle.bgzf_decompress(blob)[0] == want


def test_full_size_c2_properties(native):
    """BASELINE config C2 at full size (10 000 + 10 000 samples x 5 000 000 SNPs = 1e11 calls, 395 GB of text), output
    left on the device: size-independent properties.  Row ranges concatenate (bytes add up, the xor of the block
    CRC32s -- the checksum of checksums -- composes), the counts match the plan, and windows anywhere in the
    population decompress to the oracle's rows."""
    import bench
    from oracle import oracle
    _native, host = native
    S = bench.TOTAL_SNPS
    sex, ctl, table, orow, osamp = bench.synth_population(S, 0, window=S)
    eng = _native.Engine(0)
    eng.set_samples(sex, ctl)
    eng.set_snps(**table.device_arrays())
    eng.set_overrides(orow, osamp)
    text_bytes, _ = eng.plan(0, S)
    whole = eng.generate_device(0, S, bench.PHILOX_SEED, level=2)
    assert whole["rows"] == S and whole["calls"] == S * len(sex) == 10 ** 11
    assert whole["text_bytes"] == text_bytes > 3.9e11
    cuts = [0, 1, 777_777, S // 2, S - 3, S]
    parts = [eng.generate_device(a, b, bench.PHILOX_SEED, level=2) for a, b in zip(cuts, cuts[1:])]
    assert sum(p["text_bytes"] for p in parts) == text_bytes
    assert sum(p["calls"] for p in parts) == whole["calls"]
    x = 0
    for p in parts:
        x ^= p["crc_xor"]
    assert x == whole["crc_xor"]          # CRC32 of a block depends on its bytes only; blocks are cut on row boundaries
    # windows: first rows, the X/Y tail of the sorted list, and somewhere in the middle -- against the oracle
    from types import SimpleNamespace
    fam = [SimpleNamespace(sex=int(s), is_control=bool(c), deleterious_snps=None if c else {}, person_id=i)
           for i, (s, c) in enumerate(zip(sex, ctl))]
    for lo in (0, 2_345_678, S - 40):
        hi = lo + 40
        blob, st = eng.generate(lo, hi, bench.PHILOX_SEED, level=2)
        flat = oracle.flatten(fam, [table.snp(r) for r in range(lo, hi)])
        sel = (orow >= lo) & (orow < hi)
        flat["over_row"] = (orow[sel] - lo).astype(np.uint64)
        flat["over_sample"] = osamp[sel]
        want, _ = oracle.rows_from_flat(flat, bench.PHILOX_SEED, lo, n_threads=8)
        assert oracle.bgzf_decompress(blob)[0] == want.tobytes()


def test_randomised_shapes_soak():
    """A short run of tests/tools/fuzz_parity.py (random N, sex ratio, MAF mix incl. 1e-6 and K = 1 / 3, override density,
    pass size, level, row base) -- every case must decompress to the oracle's rows."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "tools", "fuzz_parity.py"), "16", "20261018"], capture_output=True,
                       text=True, cwd=root, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


def test_failed_calls_leave_a_usable_context(native):
    """A call that fails half way (sink raises on its second piece; caller's buffer too small) drains the pass pipeline
    before it returns, and the same context then produces the same stream as an untouched one."""
    import ctypes
    _native, _ = native
    case = synth_case(3000, 96, seed=21)
    eng = _engine(native, case, chunk=256 << 10)        # 12 KB rows: many short passes in flight
    want, _ = eng.generate(0, 96, case.seed, level=2)
    pieces = []

    def write(b):
        if len(pieces) == 1:
            raise OSError("disk full")
        pieces.append(b)

    with pytest.raises(OSError, match="disk full"):
        eng.generate_stream(0, 96, case.seed, write, level=2)
    assert len(pieces) == 1 and want.startswith(pieces[0])
    got = []
    eng.generate_stream(0, 96, case.seed, got.append, level=2)
    assert b"".join(got) == want
    small = np.empty(len(want) // 2, np.uint8)
    st = _native.Stats()
    rc = eng._lib.dnaf_generate(eng._h, 0, 96, case.seed, 2, small.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)),
                                small.nbytes, ctypes.byref(st))
    assert rc == -4 and b"too small" in eng._lib.dnaf_last_error(eng._h)
    again, _ = eng.generate(0, 96, case.seed, level=2)
    assert again == want
