#!/bin/bash
# compute-sanitizer memcheck + racecheck over tests/tools/sanitize_case.py (run on the GPU box: gpurun -- scripts/sanitize.sh TAG)
tag=${1:-r2}
mkdir -p gpurun_out
python tests/tools/sanitize_case.py > gpurun_out/sanitize_plain_$tag.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/sanitize_plain_$tag.log; exit 1; }
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 40 --log-file gpurun_out/sanitize_${tool}_$tag.log python tests/tools/sanitize_case.py > gpurun_out/sanitize_${tool}_$tag.out 2>&1
  echo "$tool exit $?"; tail -4 gpurun_out/sanitize_${tool}_$tag.log
done
