#!/bin/bash
# Run under gpurun (one GPU): parity tests, bench lines (ours + reference arm), ncu launch list and one full capture of
# the LBA solve kernel, PO launch list and timing, phase cycles, batch-size scaling, e2e split, staging experiment,
# motion-only latency.  scripts/make_profile_summary.py <tag> <round> then turns gpurun_out/ into profiles/.
# usage: scripts/gpu_profile.sh <tag>
tag=${1:-r1}
mkdir -p gpurun_out
nproc > gpurun_out/host_${tag}.txt; lscpu | grep -E "Model name|Socket|NUMA node\(s\)" >> gpurun_out/host_${tag}.txt
if [ "$2" != "notests" ]; then timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5; fi
timeout 600 python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; tail -c 3000 gpurun_out/bench_${tag}.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${tag}.json 2>> gpurun_out/bench_${tag}.err; cat gpurun_out/bench_ref_${tag}.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_a_${tag}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lba_solve -s 3 -c 1 -f -o gpurun_out/prof_lba_${tag} \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_b_${tag}.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_po_${tag}.csv \
    python scripts/po_profile.py 1 > gpurun_out/ncu_po_${tag}.log 2>&1
timeout 300 python scripts/po_profile.py 3 2>&1 | tee gpurun_out/po_${tag}.txt
timeout 300 python scripts/phase_profile.py all > gpurun_out/phase_${tag}.txt 2>&1
timeout 300 python scripts/phase_profile.py scale > gpurun_out/phase_scale_${tag}.txt 2>&1
timeout 300 python scripts/e2e_profile.py > gpurun_out/e2e_${tag}.txt 2>&1
timeout 300 python scripts/h2d_staging_probe2.py > gpurun_out/h2d_${tag}.txt 2>&1
timeout 300 python scripts/motion_only_profile.py > gpurun_out/moba_${tag}.txt 2>&1
timeout 300 python scripts/ransac_profile.py > gpurun_out/ransac_${tag}.txt 2>&1
timeout 300 python scripts/po_compare_orders.py > gpurun_out/po_orders_${tag}.txt 2>&1
timeout 300 python scripts/wide_probe.py 2>&1 | grep "===\|^ms" > gpurun_out/wide_${tag}.txt
# every other kernel family once, full metric set (the summaries go to profiles/<round>_other_kernels_ncu.csv)
timeout 900 ncu --set full --clock-control none -k regex:'^(?!.*(lba_solve_kernel|elementwise|at::|vectorized)).*$' -c 100 -f -o gpurun_out/prof_other_${tag} \
    python scripts/all_kernels_driver.py > gpurun_out/ncu_c_${tag}.log 2>&1
# gpurun merges at most 64 MiB back: keep the raw page of that capture as CSV, not the 50 MB report
ncu -i gpurun_out/prof_other_${tag}.ncu-rep --page raw --csv > gpurun_out/other_raw_${tag}.csv 2>/dev/null; rm -f gpurun_out/prof_other_${tag}.ncu-rep
./build/ubench_latency > gpurun_out/ubench_${tag}.txt 2>&1
timeout 120 python -c "from slslam_b200 import capi; print('measured fp64 peak TFLOP/s, SM MHz:', capi.measure_fp64_peak())" >> gpurun_out/ubench_${tag}.txt 2>&1
ls -la gpurun_out | tail -25
