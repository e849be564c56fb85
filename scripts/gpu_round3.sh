#!/bin/bash
# quick iteration: selected parity tests, bench, phase cycles.  usage: scripts/gpu_round3.sh <tag> [pytest -k expression]
tag=${1:-r1d}
sel=${2:-"pipeline or M_window or determinism or host_buffer or cluster_sizes or heterogeneous or large_group"}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "$sel" 2>&1 | tail -8
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; tail -c 4000 gpurun_out/bench_${tag}.json; tail -5 gpurun_out/bench_${tag}.err
timeout 300 python scripts/phase_profile.py batch0 > gpurun_out/phase_${tag}.txt 2>&1; cat gpurun_out/phase_${tag}.txt
timeout 300 python scripts/motion_only_profile.py > gpurun_out/moba_${tag}.txt 2>&1; cat gpurun_out/moba_${tag}.txt
