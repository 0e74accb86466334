#!/bin/bash
# to-file run of the re-hosted CLI on a slice of config C2: usage scripts/cli_run.sh <snps> [extra flags]
S=${1:-65536}; shift
OUT=$(mktemp -d)
time python -m dna_factory_b200.pop_factory -s 10000 -c 10000 -x $S -f 0.01 -z 2 -p tests/golden/cli_small/deleterious_config.yml \
    --outdir $OUT --seed 4242 "$@" 2>&1 | tail -25
ls -la $OUT
python - $OUT <<PY
import sys, gzip, time
t0 = time.time(); n = 0; rows = 0
with gzip.open(sys.argv[1] + "/population.vcf.gz", "rb") as f:
    while True:
        b = f.read(1 << 24)
        if not b: break
        n += len(b); rows += b.count(b"\n")
print("gzip -d check: %d bytes, %d lines, %.1f s" % (n, rows, time.time() - t0))
PY
rm -rf $OUT
