#!/usr/bin/env python
"""Where the host-buffer (one-shot) LBA solve spends its time: plan / staging / H2D / kernel / D2H."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
from slslam_b200 import capi, synth

ws = [synth.window_M(i, sigma_px=1.0, start="far") for i in range(8)]
for n in (8, 1):
    for _ in range(3):
        capi.lba_solve_batch(ws[:n], max_iters=10)
    ts, split = [], []
    for _ in range(20):
        t0 = time.perf_counter()
        ps, ss = capi.lba_solve_batch(ws[:n], max_iters=10)
        ts.append(time.perf_counter() - t0)
        split.append(capi.last_timings())
    it = sum(s["iterations"] for s in ss)
    med = {k: float(np.median([s[k] for s in split])) for k in split[0]}
    print(f"one-shot n={n}: python wall {np.median(ts)*1e3:.3f} ms -> {it/np.median(ts):.0f} it/s; C split (median) " +
          " ".join(f"{k}={v:.3f}" for k, v in med.items()), flush=True)
# raw PCIe reference: pinned 8 MB H2D and 0.5 MB D2H
h = torch.empty(7_878_656, dtype=torch.uint8).pin_memory()
d = torch.empty_like(h, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
e0.record()
for _ in range(20):
    d.copy_(h, non_blocking=True)
e1.record(); torch.cuda.synchronize()
print(f"pinned H2D 7.9 MB: {e0.elapsed_time(e1)/20:.3f} ms -> {7.878656e-3/(e0.elapsed_time(e1)/20e3):.1f} GB/s")
t0 = time.perf_counter()
for _ in range(20):
    d.copy_(h, non_blocking=True); torch.cuda.synchronize()
print(f"pinned H2D 7.9 MB incl. sync, host wall: {(time.perf_counter()-t0)/20*1e3:.3f} ms")
