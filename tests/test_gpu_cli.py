"""GPU end-to-end test of the re-hosted CLI against the reference CLI run pinned in tests/golden/cli_small."""
import gzip
import json
import os
import random

import pytest

from tests.cases import GOLDEN

pytestmark = pytest.mark.gpu


def _run(tmp_path, monkeypatch, extra):
    from dna_factory_b200 import pop_factory
    gold = os.path.join(GOLDEN, "cli_small")
    meta = json.load(open(os.path.join(gold, "meta.json")))

    class FixedDatetime(pop_factory.datetime):
        @classmethod
        def now(cls, tz=None):
            return cls(2026, 1, 1, 12, 34, 56)

    monkeypatch.setattr(pop_factory, "datetime", FixedDatetime)
    random.seed(meta["python_random_seed"])
    args = meta["args"] + ["-p", os.path.join(gold, "deleterious_config.yml"), "--outdir", str(tmp_path),
                           "--seed", str(meta["philox_seed"])] + extra
    pop_factory.main(args)
    return gold, meta


@pytest.mark.parametrize("extra", [[], ["--gpus", "1", "-n", "7"]])
def test_cli_output_directory_matches_reference(tmp_path, monkeypatch, extra):
    """Every file the reference writes, byte for byte (population.vcf.gz: its decompressed bytes)."""
    import hashlib
    from oracle import oracle
    gold, meta = _run(tmp_path, monkeypatch, extra)
    blob = (tmp_path / "population.vcf.gz").read_bytes()
    vcf = gzip.decompress(blob)
    assert hashlib.sha256(vcf).hexdigest() == meta["vcf_sha256"] and len(vcf) == meta["vcf_len"]
    text, blocks, eof = oracle.bgzf_decompress(blob)
    assert text == vcf and eof and blocks >= 3
    with gzip.open(tmp_path / "snps.json.gz", "rb") as f:
        assert f.read() == open(os.path.join(gold, "snps.json"), "rb").read()
    for name in ("deleterious.json", "population.fam", "pop_deleterious.txt"):
        assert (tmp_path / name).read_bytes() == open(os.path.join(gold, name), "rb").read(), name


def test_cli_offset_replay_matches_reference(tmp_path, monkeypatch):
    """The README's multi-run recipe (README.md:88-94): `--offset 300` replayed from snps.json.gz / deleterious.json.
    population.fam, pop_deleterious.txt and the whole VCF (header ids + rows) against the pinned reference run
    (tests/golden/cli_offset, pop_factory.py:350-351,378)."""
    import hashlib
    from dna_factory_b200 import pop_factory
    gold = os.path.join(GOLDEN, "cli_offset")
    src = os.path.join(GOLDEN, "cli_small")
    meta = json.load(open(os.path.join(gold, "meta.json")))

    class FixedDatetime(pop_factory.datetime):
        @classmethod
        def now(cls, tz=None):
            return cls(2026, 1, 1, 12, 34, 56)

    monkeypatch.setattr(pop_factory, "datetime", FixedDatetime)
    snps_gz = tmp_path / "snps_in.json.gz"
    with open(os.path.join(src, "snps.json"), "rb") as f, gzip.open(snps_gz, "wb") as g:
        g.write(f.read())
    random.seed(meta["python_random_seed"])
    out = tmp_path / "out"
    pop_factory.main(meta["args"] + ["--snps_file", str(snps_gz), "--deleterious_file", os.path.join(src, "deleterious.json"),
                                     "--outdir", str(out), "--seed", str(meta["philox_seed"])])
    vcf = gzip.decompress((out / "population.vcf.gz").read_bytes())
    assert hashlib.sha256(vcf).hexdigest() == meta["vcf_sha256"] and len(vcf) == meta["vcf_len"]
    for name in ("population.fam", "pop_deleterious.txt"):
        assert (out / name).read_bytes() == open(os.path.join(gold, name), "rb").read(), name


def test_cli_replay_from_files_reproduces_r8(tmp_path, monkeypatch):
    """--snps_file / --deleterious_file replay: string keys of deleterious.json never match the int ids, so
    no forced minors appear -- exactly what the reference does.  Checked against the oracle."""
    from dna_factory_b200 import pop_factory
    from dna_factory_b200.snp import SnpTable
    from oracle import oracle
    gold, meta = _run(tmp_path, monkeypatch, [])
    out2 = tmp_path / "replay"
    random.seed(99)
    pop_factory.main(["-s", "5", "-c", "4", "-z", "2", "--snps_file", str(tmp_path / "snps.json.gz"),
                      "--deleterious_file", str(tmp_path / "deleterious.json"), "--outdir", str(out2), "--seed", "77"])
    vcf = gzip.decompress((out2 / "population.vcf.gz").read_bytes())
    header, rows = vcf.split(b"\n", 6)[:6], vcf.split(b"\n", 6)[6]
    snps = SnpTable.read_json_gz(str(tmp_path / "snps.json.gz"))
    fam = []
    for line in (out2 / "population.fam").read_text().splitlines():
        f = line.split("\t")
        fam.append(pop_factory.SampleInfo(int(f[0]), int(f[1]), 0, 0, int(f[4]), f[5] == "1",
                                          None if f[5] == "1" else {"not-an-int-key": 1}))
    want, _ = oracle.rows(fam, snps, 77, 0)
    assert rows == want
    assert len(header) == 6


def test_cli_gpu_select(tmp_path, monkeypatch):
    """--gpu_select: snps.json.gz holds exactly the SNPs of the replay selection stream (the numpy restatement of
    SnpFactory.random_snp_tuples + sort), and the VCF rows are the oracle's rows for them."""
    from dna_factory_b200 import pop_factory, snp
    from dna_factory_b200.snp import SnpTable
    from oracle import oracle, snp_select
    gold = os.path.join(GOLDEN, "cli_small")
    random.seed(5)
    pop_factory.main(["-s", "6", "-c", "5", "-x", "400", "-f", "0.01", "-z", "2", "-p",
                      os.path.join(gold, "deleterious_config.yml"), "--outdir", str(tmp_path), "--seed", "4242",
                      "--gpu_select"])
    fac = snp.SnpFactory.init_from_cdf_file()
    t = fac.selection_tables(0.01)
    want_tab = fac.table_from_columns(snp_select.select(4242, 400, t["chrom_cdf"], t["chrom_max_pos"], t["chrom_rank"],
                                                        t["maf_cdf"]), t["start"])
    snps = SnpTable.read_json_gz(str(tmp_path / "snps.json.gz"))
    assert [(s.id, s.chromosome, s.position, s.tuples) for s in snps] == \
        [(s.id, s.chromosome, s.position, s.tuples) for s in want_tab.to_snps()]
    vcf = gzip.decompress((tmp_path / "population.vcf.gz").read_bytes())
    rows = vcf.split(b"\n", 6)[6]
    assert rows.count(b"\n") == 400


@pytest.mark.parametrize("size,control,snps", [(700, 600, 6000), (3, 2, 300), (0, 0, 500)])
def test_cli_tbi_index(tmp_path, size, control, snps):
    """--tbi: population.vcf.gz.tbi answers region queries (through an independent reader of the tabix format that
    seeks by virtual offset) exactly like a scan of the inflated file -- rows on the fused path (1300 samples),
    on the generic path (5 samples) and sites-only."""
    from dna_factory_b200 import pop_factory
    from tests import tbi_reader
    gold = os.path.join(GOLDEN, "cli_small")
    random.seed(9)
    pop_factory.main(["-s", str(size), "-c", str(control), "-x", str(snps), "-f", "0.01", "-z", "2", "-p",
                      os.path.join(gold, "deleterious_config.yml"), "--outdir", str(tmp_path), "--seed", "99",
                      "--gpu_select", "--tbi"])
    data = (tmp_path / "population.vcf.gz").read_bytes()
    tbi = tbi_reader.parse_tbi(tbi_reader.bgzf_inflate_all((tmp_path / "population.vcf.gz.tbi").read_bytes()))
    text = gzip.decompress(data)
    assert tbi_reader.bgzf_inflate_all(data) == text
    body = [ln for ln in text.splitlines(keepends=True) if not ln.startswith(b"#")]
    assert len(body) == snps
    keys = [(ln.split(b"\t", 2)[0].decode(), int(ln.split(b"\t", 2)[1])) for ln in body]
    names = []
    for c, _ in keys:
        if c not in names:
            names.append(c)
    assert tbi["names"] == names and (tbi["format"], tbi["col_seq"], tbi["col_beg"], tbi["meta"]) == (2, 1, 2, 35)
    for name, ref in zip(names, tbi["refs"]):
        assert ref["bins"][tbi_reader.META_BIN][1] == (sum(1 for c, _ in keys if c == name), 0)

    def brute(chrom, beg, end):
        return [ln for ln, (c, p) in zip(body, keys) if c == chrom and max(p - 1, 0) < end and max(p - 1, 0) + 1 > beg]

    rng = random.Random(17)
    for _ in range(40):
        chrom = rng.choice(names)
        beg = rng.randrange(0, 200_000_000)
        end = beg + rng.choice([1, 1000, 100_000, 5_000_000, 250_000_000])
        assert tbi_reader.query(data, tbi, chrom, beg, end) == brute(chrom, beg, end), (chrom, beg, end)
    for c, p in keys[::max(1, snps // 25)]:
        got = tbi_reader.query(data, tbi, c, max(p - 1, 0), max(p, 1))
        assert got == brute(c, max(p - 1, 0), max(p, 1)) and got


@pytest.mark.parametrize("ranks,level", [(2, 2), (3, 6)])
def test_cli_two_gpus_same_vcf_and_index(tmp_path, ranks, level):
    """--gpus N (contiguous SNP ranges, every rank sizes its stream and pwrites it at its offset): the
    file equals the one-GPU file byte for byte, and the index written across the ranks' block tables answers region
    queries."""
    from dna_factory_b200 import pop_factory      # on a 1-GPU box the two ranks share the device: same code path
    from tests import tbi_reader
    gold = os.path.join(GOLDEN, "cli_small")

    class Frozen(pop_factory.datetime):
        @classmethod
        def now(cls, tz=None):
            return cls(2026, 1, 1, 12, 34, 56)

    real = pop_factory.datetime
    pop_factory.datetime = Frozen
    try:
        texts = []
        for g in (1, ranks):
            out = tmp_path / ("g%d" % g)
            random.seed(7)
            pop_factory.main(["-s", "1500", "-c", "1500", "-x", "3000", "-f", "0.01", "-z", str(level), "-p",
                              os.path.join(gold, "deleterious_config.yml"), "--outdir", str(out), "--seed", "99",
                              "--gpu_select", "--tbi", "--gpus", str(g)])
            texts.append(gzip.decompress((out / "population.vcf.gz").read_bytes()))
    finally:
        pop_factory.datetime = real
    assert texts[0] == texts[1]
    data = (tmp_path / ("g%d" % ranks) / "population.vcf.gz").read_bytes()
    # (the compressed bytes may differ from the one-GPU file: a rank's literal codes are fitted to the prefix bytes of
    # ITS rows -- a slice without X rows has no 'X' literal; the sizing pass and the write pass of a rank always agree)
    tbi = tbi_reader.parse_tbi(tbi_reader.bgzf_inflate_all((tmp_path / ("g%d" % ranks) / "population.vcf.gz.tbi").read_bytes()))
    body = [ln for ln in texts[1].splitlines(keepends=True) if not ln.startswith(b"#")]
    keys = [(ln.split(b"\t", 2)[0].decode(), int(ln.split(b"\t", 2)[1])) for ln in body]
    for c, p in keys[::60]:
        want = [ln for ln, k in zip(body, keys) if k == (c, p)]
        assert tbi_reader.query(data, tbi, c, max(p - 1, 0), max(p, 1)) == want
