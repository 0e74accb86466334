"""Runs the UNMODIFIED reference CLI from oracle/_ref/ as a subprocess and reads its own timer lines
(TEST / BASELINE INFRASTRUCTURE ONLY -- used by bench.py's CPU legs; see oracle/make_ref.py).

The reference's `write_vcf_snps` is wrapped in a Timer that prints
    "Finished write_vcf_snps chunk Elapsed time: %f seconds"            (pop_factory.py:417)
once per 1 M-SNP chunk; the sum of those lines is the time of the hot path, the process wall clock adds SNP
selection, the .fam / deleterious writers and interpreter start-up.
"""
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time

from . import make_ref

HERE = os.path.dirname(os.path.abspath(__file__))
SHIMS = os.path.join(HERE, "shims")


def available():
    return make_ref.available()


def host_cores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def run(controls, cases, snps, level, procs=None, min_maf=0.01, keep_dir=None, timeout=1800):
    """`pop_factory.py -s controls -c cases -x snps -f min_maf -n procs -z level` with the reference's default
    deleterious.yml.  Returns dict(write_s, wall_s, calls, calls_per_s (hot path), procs, vcf_bytes)."""
    if not available():
        raise RuntimeError("oracle/_ref/dna_factory is missing: run `python oracle/make_ref.py` where /root/reference exists")
    procs = procs or max(1, host_cores() - 1)          # README.md:63-68: "never ... higher than cores - 1"
    out = keep_dir or tempfile.mkdtemp(prefix="dnaf_ref_")
    env = dict(os.environ)
    env["PYTHONPATH"] = SHIMS + os.pathsep + env.get("PYTHONPATH", "")
    for k in ("OMP_NUM_THREADS", "RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    cmd = [sys.executable, os.path.join(make_ref.DEST, "pop_factory.py"), "-s", str(controls), "-c", str(cases), "-x", str(snps),
           "-f", str(min_maf), "-n", str(procs), "-z", str(level), "-p", os.path.join(make_ref.DEST, "deleterious.yml"),
           "--outdir", out]
    t0 = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, cwd=make_ref.DEST, timeout=timeout)
    wall = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError("reference CLI failed: %s" % (r.stderr[-800:] or r.stdout[-800:]))
    write_s = sum(float(x) for x in re.findall(r"Finished write_vcf_snps chunk Elapsed time: ([0-9.]+) seconds", r.stdout))
    vcf = os.path.join(out, "population.vcf.gz")
    res = {"write_s": write_s, "wall_s": wall, "calls": (controls + cases) * snps, "procs": procs,
           "calls_per_s": (controls + cases) * snps / write_s if write_s > 0 else 0.0,
           "vcf_bytes": os.path.getsize(vcf) if os.path.exists(vcf) else 0, "cmd": " ".join(cmd[1:])}
    if keep_dir is None:
        shutil.rmtree(out, ignore_errors=True)
    return res
