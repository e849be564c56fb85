#!/bin/bash
# One gpurun call: parity tests, bench (ours + reference arm), batch-size scaling of the solve kernel.
# usage: scripts/gpu_round.sh <tag> [skip-tests]
tag=${1:-r1b}
mkdir -p gpurun_out
nproc; lscpu | grep -E "Model name|Socket|NUMA node\(s\)" 
if [ "$2" != "skip-tests" ]; then timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8; fi
timeout 600 python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; tail -c 4500 gpurun_out/bench_${tag}.json; tail -5 gpurun_out/bench_${tag}.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${tag}.json 2>> gpurun_out/bench_${tag}.err; cat gpurun_out/bench_ref_${tag}.json
timeout 300 python scripts/phase_profile.py scale > gpurun_out/phase_scale_${tag}.txt 2>&1; cat gpurun_out/phase_scale_${tag}.txt
