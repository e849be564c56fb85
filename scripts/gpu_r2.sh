#!/bin/bash
# usage: scripts/gpu_r2.sh <tag> [tests|notests] [bench|nobench]
tag=${1:-r2a}
mkdir -p gpurun_out
if [ "$2" != "notests" ]; then timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_${tag}.txt; fi
if [ "$3" != "nobench" ]; then
  timeout 600 python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; tail -c 3000 gpurun_out/bench_${tag}.json; tail -5 gpurun_out/bench_${tag}.err
  timeout 200 python scripts/phase_profile.py batch0 > gpurun_out/phase_${tag}.txt 2>&1; cat gpurun_out/phase_${tag}.txt
fi
