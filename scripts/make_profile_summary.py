#!/usr/bin/env python
"""Turn gpurun_out/ ncu artefacts into the tracked summaries under profiles/.

usage: make_profile_summary.py <tag> <round-name>
  reads  gpurun_out/prof_lba_<tag>.ncu-rep, gpurun_out/launches_<tag>.csv, gpurun_out/launches_po_<tag>.csv,
         gpurun_out/bench_<tag>.json, gpurun_out/bench_ref_<tag>.json, gpurun_out/po_<tag>.txt
  writes profiles/<round>_*.{csv,txt,json}
"""
import csv
import io
import os
import shutil
import subprocess
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, rnd = sys.argv[1], sys.argv[2]
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__sass_thread_inst_executed_op_dfma_pred_on.sum", "sm__sass_thread_inst_executed_op_dfma_pred_on.sum.peak_sustained",
        "sm__sass_thread_inst_executed_op_dadd_pred_on.sum", "sm__sass_thread_inst_executed_op_dmul_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed", "smsp__cycles_elapsed.avg", "smsp__cycles_elapsed.max",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "memory_l1_wavefronts_shared", "memory_l1_wavefronts_shared_ideal",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "sm__cycles_elapsed.avg", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__cluster_size", "launch__cluster_max_active", "launch__shared_mem_per_block_dynamic",
        "launch__waves_per_multiprocessor", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warp_latency_per_inst_issued.ratio", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]

rep = os.path.join(G, f"prof_lba_{tag}.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(os.path.join(P, f"{rnd}_lba_solve_kernel_ncu_raw.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(rows) - 2)])
        name_col = hdr.index("Kernel Name")
        w.writerow(["Kernel Name", ""] + [r[name_col] for r in rows[2:]])
        for i, h in enumerate(hdr):
            if h in KEEP or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") \
                    or h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued"):
                w.writerow([h, units[i]] + [r[i] for r in rows[2:]])
    by = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_by_line.py"), rep,
                         os.path.join(ROOT, "slslam_b200", "libslslam_b200.so"), "lba_solve", "60"], capture_output=True, text=True).stdout
    open(os.path.join(P, f"{rnd}_lba_solve_kernel_by_source_line.txt"), "w").write(
        "# ncu --set full --import-source on, samples and instructions aggregated by CUDA source line (scripts/ncu_by_line.py)\n" + by)

for src, dst in ((f"launches_{tag}.csv", f"{rnd}_launches_bench_lba.csv"), (f"launches_po_{tag}.csv", f"{rnd}_launches_po_solve.csv"),
                 (f"bench_{tag}.json", f"{rnd}_bench_ours.json"), (f"bench_ref_{tag}.json", f"{rnd}_bench_reference_arm.json"),
                 (f"po_{tag}.txt", f"{rnd}_po_solve_timing.txt"), (f"phase_{tag}.txt", f"{rnd}_lba_phase_cycles.txt"),
                 (f"e2e_{tag}.txt", f"{rnd}_lba_e2e_split.txt"), (f"h2d_{tag}.txt", f"{rnd}_h2d_staging.txt"),
                 (f"phase_scale_{tag}.txt", f"{rnd}_lba_batch_size_scaling.txt"), (f"moba_{tag}.txt", f"{rnd}_motion_only_ba.txt"),
                 (f"host_{tag}.txt", f"{rnd}_host_cpu.txt"), (f"ransac_{tag}.txt", f"{rnd}_ransac_scoring.txt"),
                 (f"po_orders_{tag}.txt", f"{rnd}_po_level_vs_column_order.txt"), (f"wide_{tag}.txt", f"{rnd}_wide_kernel_timing.txt")):
    if os.path.exists(os.path.join(G, src)):
        shutil.copy(os.path.join(G, src), os.path.join(P, dst))

for src, dst in ((f"ubench_{tag}.txt", f"{rnd}_latency_and_fp64_peak.txt"),):
    if os.path.exists(os.path.join(G, src)):
        shutil.copy(os.path.join(G, src), os.path.join(P, dst))

# the other kernel families: one row per launch of the `ncu --set full` capture of scripts/all_kernels_driver.py
rep2 = os.path.join(G, f"other_raw_{tag}.csv")
if os.path.exists(rep2):
    rows = list(csv.reader(open(rep2)))
    if len(rows) > 2:
        hdr, units = rows[0], rows[1]
        name_col = hdr.index("Kernel Name")
        cols = [i for i, h in enumerate(hdr) if h in KEEP or h in ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
                                                                   "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
                                                                   "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
                                                                   "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio")]
        with open(os.path.join(P, f"{rnd}_other_kernels_ncu.csv"), "w") as f:
            w = csv.writer(f)
            w.writerow(["# ncu --set full --clock-control none, scripts/all_kernels_driver.py: one row per launch"])
            w.writerow(["kernel"] + [hdr[i] + (" [" + units[i] + "]" if units[i] else "") for i in cols])
            for r in rows[2:]:
                w.writerow([r[name_col].split("(")[0]] + [r[i] for i in cols])

# per-kernel totals of the PO launch list
po = os.path.join(G, f"launches_po_{tag}.csv")
if os.path.exists(po):
    tot, cnt = defaultdict(float), defaultdict(int)
    rows = [r for r in csv.reader(open(po)) if len(r) > 10]
    if rows:
        hdr = rows[0]
        kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
        for r in rows[1:]:
            try:
                name = r[kn].split("(")[0]
                tot[name] += float(r[mv].replace(",", "")); cnt[name] += 1
            except Exception:
                pass
        all_ns = sum(tot.values())
        with open(os.path.join(P, f"{rnd}_po_solve_kernel_shares.txt"), "w") as f:
            f.write("# ncu --metrics gpu__time_duration.sum --clock-control none, one slslam_po_solve (261 poses, 785 edges); shares, not absolutes\n")
            for k in sorted(tot, key=lambda k: -tot[k]):
                f.write(f"{k:40s} launches {cnt[k]:5d}  total {tot[k]/1e3:10.1f} us  share {100*tot[k]/all_ns:5.1f} %\n")
print("profiles/:", sorted(os.listdir(P)))
