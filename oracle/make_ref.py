"""Materialises the UNMODIFIED reference into oracle/_ref/ (TEST / BASELINE INFRASTRUCTURE ONLY).

    python oracle/make_ref.py            # copies /root/reference's pop_factory path into oracle/_ref/dna_factory/

The reference is pure Python (no build step): the "recipe" is a copy of the files its population-generation path
imports -- pop_factory.py, definitions.py, common/, snp_freq_cdf.csv, deleterious.yml -- byte for byte, with a
MANIFEST of their sha256 sums.  oracle/_ref/ is git-ignored (no reference source enters this repository's history)
but NOT gpurun-ignored, so the copy travels to the GPU box with the snapshot, where bench.py's CPU legs
(`cpu_baseline`, `--impl reference`) run it as `python pop_factory.py -n <cores-1> ...` under the two import shims
of oracle/shims/ (biopython's BgzfWriter restated on zlib, a dummy sqlalchemy).  Nothing in the product
package reads oracle/_ref/.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_DIR = os.environ.get("DNAF_REFERENCE_DIR", "/root/reference")
DEST = os.path.join(HERE, "_ref", "dna_factory")
FILES = ["pop_factory.py", "definitions.py", "snp_freq_cdf.csv", "deleterious.yml", "common/__init__.py", "common/db.py",
         "common/snp.py", "common/synchro.py", "common/timer.py", "LICENSE"]


def available():
    """The copy exists (this container after build(), or the GPU box that received the snapshot)."""
    return os.path.exists(os.path.join(DEST, "pop_factory.py"))


def materialise(force=False):
    if not os.path.exists(os.path.join(REFERENCE_DIR, "pop_factory.py")):
        return available()          # no reference tree here (the GPU box): use what travelled
    manifest = {}
    for rel in FILES:
        src = os.path.join(REFERENCE_DIR, rel)
        if not os.path.exists(src):
            continue
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        with open(src, "rb") as f:
            data = f.read()
        manifest[rel] = hashlib.sha256(data).hexdigest()
        if force or not os.path.exists(dst) or open(dst, "rb").read() != data:
            shutil.copyfile(src, dst)
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REFERENCE_DIR, "sha256": manifest}, f, indent=1, sort_keys=True)
    return True


if __name__ == "__main__":
    ok = materialise(force="--force" in sys.argv)
    print("oracle/_ref/dna_factory: %s" % ("ready" if ok else "reference tree not available"))
