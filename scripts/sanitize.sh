#!/bin/bash
# compute-sanitizer over small invocations of every kernel (under gpurun).  usage: scripts/sanitize.sh <tag>
tag=${1:-r1}
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  echo "== $tool" >> gpurun_out/sanitize_${tag}.txt
  timeout 600 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize_case.py all 2>&1 | grep -E "^lba|^moba|^ransac|^po|^map|ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard|Invalid" | head -30 >> gpurun_out/sanitize_${tag}.txt
done
cat gpurun_out/sanitize_${tag}.txt
