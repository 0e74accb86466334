#!/usr/bin/env python
"""Generates the committed golden fixtures by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

For every case it builds the reference's own SNPTuples / SampleInfo objects, runs the reference's
worker loop `PopulationFactory.queue_vcf_snps` (pop_factory.py:471-513) with `numpy.random.rand`
patched to the replay Philox stream (oracle/philox_np.py), and stores

    <case>.json      inputs (seed, row_begin, samples, snps) + sha256/length of the expected text
    <case>.rows.gz   the expected row text, gzip -9

`cli_small/` is a full end-to-end run of the reference CLI (`pop_factory.main`) with the clock,
`random` and numpy seeds pinned and the forked worker's `rand` patched to the same stream, so
every output file of the reference (snps.json.gz, deleterious.json, population.fam,
pop_deleterious.txt, population.vcf.gz) is pinned byte for byte (VCF: decompressed bytes).
"""
import gzip
import hashlib
import json
import os
import random
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import philox_np, ref_harness  # noqa: E402

CHROMS = ['1', '2', '3', '4', '5', '6', '7', '8', '9', '10', '11', '12', '13', '14', '15', '16', '17', '18', '19',
          '20', '21', '22', 'X', 'Y']


def dump_case(name, seed, row_begin, samples, snps, text):
    meta = {
        "seed": seed,
        "row_begin": row_begin,
        "samples": [
            {"family_id": s.family_id, "person_id": s.person_id, "sex": s.sex, "is_control": bool(s.is_control),
             "deleterious": None if s.deleterious_snps is None else
             [[k, "int" if isinstance(k, int) else "str", w] for k, w in s.deleterious_snps.items()]}
            for s in samples],
        "snps": [{"id": s.id, "chromosome": s.chromosome, "position": s.position,
                  "tuples": [[t[0], t[1]] for t in s.tuples]} for s in snps],
        "text_len": len(text),
        "text_sha256": hashlib.sha256(text).hexdigest(),
    }
    with open(os.path.join(HERE, name + ".json"), "w") as f:
        json.dump(meta, f, separators=(",", ":"))
    with open(os.path.join(HERE, name + ".rows.gz"), "wb") as raw:
        with gzip.GzipFile(filename="", mode="wb", fileobj=raw, compresslevel=9, mtime=0) as f:
            f.write(text)
    print("%-14s rows=%d samples=%d text=%d bytes" % (name, len(snps), len(samples), len(text)))


def build_samples(rs, n_control, n_case, male_odds, snps, n_del, str_keys=False):
    samples = []
    cand = [s.id for s in snps]
    for i in range(n_control + n_case):
        is_control = i < n_control
        sex = 1 if rs.rand() <= male_odds else 2
        d = None
        if not is_control:
            ids = rs.choice(cand, size=min(n_del, len(cand)), replace=False)
            d = {}
            for x in ids:
                k = int(x)
                d[str(k) if str_keys else k] = 0.5
        pid = (100001 + i) if is_control else (500001 + i - n_control)
        samples.append(ref_harness.make_sample(i, pid, sex, is_control, d))
    return samples


def build_snps(rs, n, chroms, exotic=True):
    snps = []
    for i in range(n):
        c = chroms[rs.randint(len(chroms))]
        pos = int(rs.rand() * 2.4e8)
        nts = ["A", "T", "C", "G"]
        rs.shuffle(nts)
        maf = [0.01, 0.015, 0.05, 0.1, 0.25, 0.33, 0.495, 0.005][rs.randint(8)]
        kind = rs.randint(12) if exotic else 0
        if kind == 9:      # three alleles (DB-mode shape, pop_factory.py:313-331)
            tup = [(nts[0], 0.55), (nts[1], 0.8), (nts[2], 1.0)]
        elif kind == 10:   # four alleles
            tup = [(nts[0], 0.4), (nts[1], 0.7), (nts[2], 0.9), (nts[3], 1.0)]
        elif kind == 11:   # single allele
            tup = [(nts[0], 1.0)]
        else:
            tup = [(nts[0], 1 - maf), (nts[1], 1.0)]
        snps.append(ref_harness.make_snp(i + 1, c, pos, tup))
    snps.sort(key=lambda x: (x.chromosome, x.position))   # pop_factory.py:245
    return snps


def case_mixed64():
    rs = np.random.RandomState(64)
    seed, row_begin = 0x5EED000000000001, 0
    snps = build_snps(rs, 180, ['1', '10', '2', '22', '9', 'X', 'X', 'Y', 'Y', 'MT'])
    # u exactly equal to a threshold (tests the inclusive >=) and just above it
    n_all = 128
    for r, j, delta in ((3, 5, 0), (4, 6, -1), (5, 64, 0), (6, 127, -1)):
        U = int(philox_np.uniform_bits(seed, row_begin + r, n_all)[j])
        s = snps[r]
        s.chromosome = '1'
        s.tuples = [(s.tuples[0][0], (U + delta) * 2.0 ** -32), ("G" if s.tuples[0][0] != "G" else "A", 1.0)]
    snps.sort(key=lambda x: (x.chromosome, x.position))
    samples = build_samples(rs, 40, 24, 0.5, snps, 9)
    text = ref_harness.reference_rows(samples, snps, seed, row_begin)
    dump_case("mixed64", seed, row_begin, samples, snps, text)


def case_r8_strkeys():
    rs = np.random.RandomState(64)
    seed, row_begin = 0x5EED000000000001, 0
    snps = build_snps(rs, 120, ['1', '2', 'X', 'Y'])
    samples = build_samples(rs, 10, 22, 0.5, snps, 9, str_keys=True)
    text = ref_harness.reference_rows(samples, snps, seed, row_begin)
    dump_case("r8_strkeys", seed, row_begin, samples, snps, text)


def case_n0():
    rs = np.random.RandomState(7)
    snps = build_snps(rs, 25, CHROMS)
    text = ref_harness.reference_rows([], snps, 99, 0)
    dump_case("n0", 99, 0, [], snps, text)


def case_wide():
    rs = np.random.RandomState(11)
    seed, row_begin = 20260101, (1 << 32) + 5      # exercises the high half of the row counter
    snps = build_snps(rs, 14, ['1', '7', 'X', 'Y', 'MT', '3'], exotic=True)
    samples = build_samples(rs, 1100, 407, 0.5, snps, 4)
    text = ref_harness.reference_rows(samples, snps, seed, row_begin)
    dump_case("wide", seed, row_begin, samples, snps, text)


def case_single_sex():
    rs = np.random.RandomState(5)
    snps = build_snps(rs, 30, ['X', 'Y', '5'], exotic=False)
    males = build_samples(rs, 17, 0, 1.0, snps, 0)
    text = ref_harness.reference_rows(males, snps, 1, 1000)
    dump_case("all_male", 1, 1000, males, snps, text)
    females = build_samples(rs, 3, 30, -1.0, snps, 3)
    text = ref_harness.reference_rows(females, snps, 2, 7)
    dump_case("all_female", 2, 7, females, snps, text)


def case_snp_select():
    """SnpFactory.random_snp_tuples + the sort of pop_factory.py:245, run by the reference itself with its random
    sources patched to the selection stream (oracle/ref_harness.reference_snp_selection)."""
    out = []
    for size, min_maf, seed in ((3000, 0.01, 0x5EED000000000001), (1500, 0.16, 42), (1, 0.005, 3)):
        snps = ref_harness.reference_snp_selection(size, min_maf, seed)
        out.append({"size": size, "min_maf": min_maf, "seed": seed,
                    "snps": [[int(s.id), str(s.chromosome), int(s.position), str(s.tuples[0][0]), float(s.tuples[0][1]),
                              str(s.tuples[1][0]), float(s.tuples[1][1])] for s in snps]})
    with gzip.GzipFile(os.path.join(HERE, "snp_select.json.gz"), "wb", compresslevel=9, mtime=0) as f:
        f.write(json.dumps(out, separators=(",", ":")).encode())
    print("snp_select     %s SNPs" % [c["size"] for c in out])


def case_cli_small():
    """Full reference CLI run with every source of nondeterminism pinned."""
    ref = ref_harness.load()
    out = os.path.join(HERE, "cli_small")
    tmp = "/tmp/dnaf_golden_cli"
    shutil.rmtree(tmp, ignore_errors=True)
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    philox_seed = 123456          # the CLI mirror derives its Philox seed from the same HHMMSS value
    parent = os.getpid()
    state = {"k": 0}
    real_rand = ref.numpy.random.rand

    class FixedDatetime(ref.datetime):
        @classmethod
        def now(cls, tz=None):
            return cls(2026, 1, 1, 12, 34, 56)

    def rand(*shape):
        if os.getpid() == parent:
            return real_rand(*shape)            # sexes in generate_fam_file (pop_factory.py:352)
        u = philox_np.uniforms(philox_seed, state["k"], shape[0])   # forked worker: row k of the only chunk
        state["k"] += 1
        return u

    # the deleterious config is a run input: keep a re-serialised copy of it beside the outputs
    import yaml
    with open(os.path.join(ref_harness.REFERENCE_DIR, "deleterious.yml")) as f:
        cfg = yaml.safe_load(f)
    yml = os.path.join(out, "deleterious_config.yml")
    with open(yml, "w") as f:
        yaml.safe_dump(cfg, f, sort_keys=False)
    base_args = ["-s", "7", "-c", "9", "-x", "400", "-f", "0.01", "-n", "1", "-z", "6", "-m", "0.5"]
    args = base_args + ["-p", yml, "--outdir", tmp]
    real_dt = ref.datetime
    ref.datetime = FixedDatetime
    ref.numpy.random.rand = rand
    random.seed(4242)
    try:
        ref.main(args)
    finally:
        ref.datetime = real_dt
        ref.numpy.random.rand = real_rand
    with gzip.open(os.path.join(tmp, "population.vcf.gz"), "rb") as f:
        vcf = f.read()
    with gzip.open(os.path.join(tmp, "snps.json.gz"), "rb") as f:
        snps_txt = f.read()
    for name in ("deleterious.json", "population.fam", "pop_deleterious.txt"):
        shutil.copy(os.path.join(tmp, name), os.path.join(out, name))
    with open(os.path.join(out, "snps.json"), "wb") as f:
        f.write(snps_txt)
    with open(os.path.join(out, "population.vcf.rows.gz"), "wb") as raw:
        with gzip.GzipFile(filename="", mode="wb", fileobj=raw, compresslevel=9, mtime=0) as f:
            f.write(vcf)
    with open(os.path.join(out, "meta.json"), "w") as f:
        json.dump({"args": base_args, "numpy_seed": 123456, "python_random_seed": 4242,
                   "philox_seed": philox_seed, "filedate": "20260101 12:34",
                   "vcf_sha256": hashlib.sha256(vcf).hexdigest(), "vcf_len": len(vcf)}, f)
    print("cli_small      vcf=%d bytes, %d lines" % (len(vcf), vcf.count(b"\n")))


def case_cli_offset():
    """The README's multi-run recipe (README.md:88-94): a replay from snps.json.gz / deleterious.json with `--offset 300`
    (pop_factory.py:350-351,378: control ids 100000+offset, case ids 500000+offset, family ids i+1+2*offset).  Pins
    population.fam, pop_deleterious.txt and the VCF (header sample ids + rows) of the UNMODIFIED reference."""
    ref = ref_harness.load()
    out = os.path.join(HERE, "cli_offset")
    src = os.path.join(HERE, "cli_small")
    tmp = "/tmp/dnaf_golden_cli_offset"
    shutil.rmtree(tmp, ignore_errors=True)
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    os.makedirs(tmp)
    philox_seed = 123456
    parent = os.getpid()
    state = {"k": 0}
    real_rand = ref.numpy.random.rand

    class FixedDatetime(ref.datetime):
        @classmethod
        def now(cls, tz=None):
            return cls(2026, 1, 1, 12, 34, 56)

    def rand(*shape):
        if os.getpid() == parent:
            return real_rand(*shape)
        u = philox_np.uniforms(philox_seed, state["k"], shape[0])
        state["k"] += 1
        return u

    snps_gz = os.path.join(tmp, "snps_in.json.gz")
    with open(os.path.join(src, "snps.json"), "rb") as f, gzip.open(snps_gz, "wb") as g:
        g.write(f.read())
    base_args = ["-s", "6", "-c", "5", "-z", "2", "-m", "0.5", "-n", "1", "--offset", "300"]
    args = base_args + ["--snps_file", snps_gz, "--deleterious_file", os.path.join(src, "deleterious.json"), "--outdir",
                        os.path.join(tmp, "out")]
    real_dt = ref.datetime
    ref.datetime = FixedDatetime
    ref.numpy.random.rand = rand
    random.seed(777)
    try:
        ref.main(args)
    finally:
        ref.datetime = real_dt
        ref.numpy.random.rand = real_rand
    with gzip.open(os.path.join(tmp, "out", "population.vcf.gz"), "rb") as f:
        vcf = f.read()
    for name in ("population.fam", "pop_deleterious.txt"):
        shutil.copy(os.path.join(tmp, "out", name), os.path.join(out, name))
    with open(os.path.join(out, "population.vcf.rows.gz"), "wb") as raw:
        with gzip.GzipFile(filename="", mode="wb", fileobj=raw, compresslevel=9, mtime=0) as f:
            f.write(vcf)
    with open(os.path.join(out, "meta.json"), "w") as f:
        json.dump({"args": base_args, "numpy_seed": 123456, "python_random_seed": 777, "philox_seed": philox_seed,
                   "filedate": "20260101 12:34", "inputs": "cli_small/snps.json (gzipped), cli_small/deleterious.json",
                   "vcf_sha256": hashlib.sha256(vcf).hexdigest(), "vcf_len": len(vcf)}, f)
    print("cli_offset     vcf=%d bytes, %d lines" % (len(vcf), vcf.count(b"\n")))


if __name__ == "__main__":
    if sys.argv[1:] == ["cli_offset"]:
        case_cli_offset()
        sys.exit(0)
    case_snp_select()
    case_mixed64()
    case_r8_strkeys()
    case_n0()
    case_wide()
    case_single_sex()
    case_cli_small()
    case_cli_offset()
