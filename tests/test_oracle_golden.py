"""CPU tests: the oracle restatement against the reference's golden vectors and known-answer tests."""
import gzip
import zlib

import numpy as np
import pytest

from oracle import oracle, philox_np
from tests.cases import ROW_CASES, load_case, synth_case

# Random123 known-answer vectors for philox4x32-10 (kat_vectors)
PHILOX_KATS = [
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
]


@pytest.mark.parametrize("ctr,key,out", PHILOX_KATS)
def test_philox_known_answers(ctr, key, out):
    assert oracle.philox4x32_10(ctr, key) == out
    assert [int(x) for x in philox_np.philox4x32_10(np.array(ctr, dtype=np.uint64), key)] == out


def test_uniform_stream_c_matches_numpy():
    for seed, row, n in [(1, 0, 1), (0x5EED000000000001, 3, 64), (20260101, (1 << 32) + 5, 3014), (7, 9, 0)]:
        a = oracle.uniform_bits(seed, row, n)
        b = philox_np.uniform_bits(seed, row, n)
        assert np.array_equal(a, b)
        assert np.array_equal(oracle.uniforms(seed, row, n), philox_np.uniforms(seed, row, n))


def test_pick_allele_index_reference_kat():
    # test/unit/pop_factory_test.py:24-28 -- CDF [G .70, A .90, T 1.0]
    cum = [0.70, 0.90, 1.0]
    assert oracle.pick_allele_index(cum, 0.95) == 2
    assert oracle.pick_allele_index(cum, 0.4) == 0
    assert oracle.pick_allele_index(cum, 0.70) == 0      # inclusive >=
    assert oracle.pick_allele_index([0.5], 0.75) is None  # falls off the end -> None in the reference


@pytest.mark.parametrize("name", ROW_CASES)
def test_oracle_rows_match_reference_golden(name):
    case = load_case(name)
    text, row_off = oracle.rows(case.samples, case.snps, case.seed, case.row_begin)
    assert text == case.text
    assert int(row_off[-1]) == len(case.text)
    assert text.count(b"\n") == len(case.snps)


def test_oracle_rows_threaded_equals_serial():
    case = synth_case(300, 40, seed=3)
    a, _ = oracle.rows(case.samples, case.snps, case.seed, 5, n_threads=1)
    b, _ = oracle.rows(case.samples, case.snps, case.seed, 5, n_threads=4)
    assert a == b


def test_r8_string_keys_do_not_fire():
    # deleterious.json replay keeps string keys, snp.id is int -> overrides never apply
    case = load_case("r8_strkeys")
    orow, _ = oracle.override_pairs(case.samples, case.snps)
    assert len(orow) == 0
    case2 = load_case("mixed64")
    orow2, _ = oracle.override_pairs(case2.samples, case2.snps)
    assert len(orow2) > 0


@pytest.mark.parametrize("level", [1, 6, 9])
def test_oracle_bgzf_roundtrip(level):
    case = load_case("wide")
    blob = oracle.bgzf(case.text, level=level)
    assert gzip.decompress(blob) == case.text
    text, blocks, eof = oracle.bgzf_decompress(blob)
    assert text == case.text and eof
    assert blocks == (len(case.text) + 65535) // 65536 + 1
    # Biopython framing: every full block holds exactly 65536 bytes
    assert blob[:4] == b"\x1f\x8b\x08\x04" and blob[-28:] == bytes.fromhex(
        "1f8b08040000000000ff0600424302001b0003000000000000000000")


def test_oracle_bgzf_matches_shim_writer(tmp_path):
    # the C port and the Bio.bgzf shim (what the reference ran with) produce identical bytes (same libz)
    import sys, os
    shims = os.path.join(os.path.dirname(oracle.__file__), "shims")
    sys.path.insert(0, shims)
    try:
        from Bio import bgzf
    finally:
        sys.path.remove(shims)
    case = load_case("wide")
    p = tmp_path / "x.gz"
    with bgzf.BgzfWriter(filename=str(p), mode="wt+", compresslevel=2) as f:
        f.write(case.text.decode("latin-1"))
    assert p.read_bytes() == oracle.bgzf(case.text, level=2)
    for k in [k for k in sys.modules if k == "Bio" or k.startswith("Bio.")]:
        del sys.modules[k]


def test_reference_live_crosscheck():
    """Where the reference tree is mounted, re-run it live against the oracle on a fresh random case."""
    from oracle import ref_harness
    if not ref_harness.available():
        pytest.skip("reference tree not mounted (GPU box)")
    case = synth_case(70, 25, seed=9)
    snps = [ref_harness.make_snp(s.id, s.chromosome, s.position, s.tuples) for s in case.snps]
    fam = [ref_harness.make_sample(i, s.person_id, s.sex, s.is_control, s.deleterious_snps)
           for i, s in enumerate(case.samples)]
    want = ref_harness.reference_rows(fam, snps, case.seed, 11)
    got, _ = oracle.rows(case.samples, case.snps, case.seed, 11)
    assert got == want
