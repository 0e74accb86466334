#!/bin/bash
# to-file run of the re-hosted CLI on a slice of config C4 (100 000 samples), 1 vs N GPUs:
#   scripts/cli_multi.sh <snps> <gpus> <level> [outdir root, default /dev/shm]
S=${1:-65536}; G=${2:-8}; Z=${3:-6}; ROOT=${4:-/dev/shm}
for g in 1 $G; do
  OUT=$(mktemp -d -p $ROOT)
  /usr/bin/time -f "gpus=$g wall %e s" python -m dna_factory_b200.pop_factory -s 50000 -c 50000 -x $S -f 0.01 -z $Z \
      -p tests/golden/cli_small/deleterious_config.yml --outdir $OUT --seed 4242 --gpu_select --gpus $g 2>&1 | grep -E "write_vcf_snps|Finished Generating|wall"
  ls -l $OUT/population.vcf.gz | awk '{print "population.vcf.gz bytes", $5}'
  md5sum $OUT/population.vcf.gz | cut -c1-32
  rm -rf $OUT
done
