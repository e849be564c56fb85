#!/bin/bash
# One gpurun call: selected parity tests, bench, one full ncu capture of the solve kernel, phase cycles.
# usage: scripts/gpu_round2.sh <tag> [pytest -k expression]
tag=${1:-r1c}
sel=${2:-"pipeline or M_window or determinism or host_buffer or cluster_sizes or heterogeneous"}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "$sel" 2>&1 | tail -8
timeout 600 python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; tail -c 4500 gpurun_out/bench_${tag}.json; tail -5 gpurun_out/bench_${tag}.err
timeout 300 python scripts/phase_profile.py batch0 > gpurun_out/phase_${tag}.txt 2>&1; cat gpurun_out/phase_${tag}.txt
cp slslam_b200/libslslam_b200.so gpurun_out/lib_${tag}.so
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lba_solve -s 3 -c 1 -f -o gpurun_out/prof_lba_${tag} \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_${tag}.log 2>&1; tail -3 gpurun_out/ncu_${tag}.log
ls -la gpurun_out | tail -8
