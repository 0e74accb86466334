"""Runs the UNMODIFIED reference (/root/reference) in-process as the ground truth (TEST INFRASTRUCTURE ONLY).

Works only in the build container, where /root/reference is mounted; it is used by
tests/golden/make_golden.py to produce the committed fixtures and by the (auto-skipping)
cross-check tests.  Nothing that runs on the GPU box imports this module.

Import needs two shims because biopython and sqlalchemy are not installed:
oracle/shims/Bio/bgzf.py and oracle/shims/sqlalchemy/__init__.py.
"""
import os
import sys

import numpy as np

from . import philox_np

REFERENCE_DIR = os.environ.get("DNAF_REFERENCE_DIR", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
_mod = None


def available():
    return os.path.exists(os.path.join(REFERENCE_DIR, "pop_factory.py"))


def load():
    """Import the reference's pop_factory module (cached)."""
    global _mod
    if _mod is None:
        if not available():
            raise RuntimeError("reference tree not present at %s" % REFERENCE_DIR)
        saved = list(sys.path)
        stash = {k: sys.modules.pop(k) for k in list(sys.modules)
                 if k in ("pop_factory", "common", "definitions", "Bio", "sqlalchemy")
                 or k.startswith(("common.", "Bio.", "sqlalchemy."))}
        try:
            sys.path[:0] = [_SHIMS, REFERENCE_DIR]
            import pop_factory as ref  # noqa: the reference module
            _mod = ref
        finally:
            sys.path[:] = saved
            for k in ("pop_factory", "common", "definitions", "Bio", "sqlalchemy"):
                sys.modules.pop(k, None)
            for k in list(sys.modules):
                if k.startswith(("common.", "Bio.", "sqlalchemy.")):
                    sys.modules.pop(k, None)
            sys.modules.update(stash)
    return _mod


class _ListPipe:
    def __init__(self):
        self.items = []

    def send(self, x):
        self.items.append(x)

    def close(self):
        pass


def reference_rows(fam_data, snps, seed, row_begin=0):
    """Text rows from the reference's own worker loop (pop_factory.py:471-513).

    `numpy.random.rand` is patched so that its k-th call returns the replay uniforms of global
    row `row_begin + k`; `time.sleep` is patched out.  fam_data / snps must be the reference's own
    SampleInfo / SNPTuples objects.
    """
    ref = load()
    state = {"k": 0}

    def fake_rand(n):
        u = philox_np.uniforms(seed, row_begin + state["k"], n)
        state["k"] += 1
        return u

    real_rand, real_sleep = ref.numpy.random.rand, ref.time.sleep
    ref.numpy.random.rand = fake_rand
    ref.time.sleep = lambda s: None
    try:
        pf = ref.PopulationFactory(num_processes=1, generate_snps=True, output_path="/tmp/unused")
        pipe = _ListPipe()
        pf.queue_vcf_snps(fam_data, list(enumerate(snps, start=1)), pipe)
    finally:
        ref.numpy.random.rand, ref.time.sleep = real_rand, real_sleep
    assert pipe.items[-1] == "DONE"
    lines = [x[1] for x in pipe.items[:-1]]
    assert [x[0] for x in pipe.items[:-1]] == list(range(1, len(snps) + 1))
    return "".join(lines).encode("latin-1")


def make_snp(snp_id, chromosome, position, tuples):
    ref = load()
    s = ref.SNPTuples(snp_id, chromosome, position)
    for nt, c in tuples:
        s.add_tuple(nt, c)
    return s


def make_sample(index, person_id, sex, is_control, deleterious_snps, offset=0):
    ref = load()
    return ref.SampleInfo(index + 1 + offset * 2, person_id, 0, 0, sex, is_control, deleterious_snps)


def reference_snp_selection(size, min_maf, seed):
    """The reference's own SnpFactory.random_snp_tuples (pop_factory.py:160-193) and sort (:245), with its three
    random sources patched to the replay stream of oracle/snp_select.py: numpy.random.choice (chromosomes, MAFs --
    numpy's own cdf/searchsorted algorithm -- and reference nucleotides), numpy.random.random (positions) and
    random.choice (alternate nucleotide, once per SNP in draw order).  Returns the reference's SNPTuples list."""
    from . import snp_select
    ref = load()
    u = snp_select.uniforms(seed, size)
    calls = {"choice": 0, "alt": 0}

    def fake_choice(a, size=None, p=None):
        k = calls["choice"]
        calls["choice"] += 1
        a = np.asarray(a)
        if p is not None:
            cdf = np.asarray(p, dtype=np.float64).cumsum()
            cdf /= cdf[-1]
            return a[cdf.searchsorted(u["chrom" if k == 0 else "maf"], side="right")]
        return a[(u["ref"] * len(a)).astype(np.int64)]

    def fake_random(n):
        return u["pos"]

    def fake_py_choice(seq):
        i = calls["alt"]
        calls["alt"] += 1
        return seq[int(u["alt"][i] * len(seq))]

    saved = (ref.numpy.random.choice, ref.numpy.random.random, ref.random.choice)
    ref.numpy.random.choice, ref.numpy.random.random, ref.random.choice = fake_choice, fake_random, fake_py_choice
    try:
        snps = ref.SnpFactory.init_from_cdf_file().random_snp_tuples(size, min_maf=min_maf)
    finally:
        ref.numpy.random.choice, ref.numpy.random.random, ref.random.choice = saved
    assert calls["choice"] == 3 and calls["alt"] == size
    snps.sort(key=lambda x: (x.chromosome, x.position))                    # pop_factory.py:245
    return snps
