#!/usr/bin/env python
"""Per-phase SM cycles of the LBA kernel (CTA 0 of window 0) and event-timed duration for several cluster sizes."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from slslam_b200 import capi, synth

def run(windows, cs, reps=10):
    try:
        b = capi.LbaBatch(windows, cluster_size=cs, max_iters=10)
    except capi.SlslamError as e:
        print(f"nwin={len(windows)} CS={cs}: {e}")
        return
    st = torch.cuda.current_stream(); sp = st.cuda_stream
    for _ in range(3): b.solve(sp)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps): b.solve(sp)
    e1.record(st); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    _, ss = b.download(sp)
    it = sum(s["iterations"] for s in ss)
    ph = b.phase_cycles(0, sp)
    info = b.info(); pc = b.plan_cycles(0); b.close()
    n0 = ss[0]["iterations"]
    print(f"nwin={len(windows)} CS={info['cluster_size']} act={info['max_active_clusters']} zsm={info['z_in_smem']} smem={info['smem_bytes_per_cta']} ms={ms:.3f} iters={it} "
          f"-> {it/ms*1e3:.0f} it/s, {ms*1e3/ n0:.1f} us/iter(win0)")
    if pc:
        print("   plan kernel, cumulative cycles: " + " ".join(f"{k}={v}" for k, v in pc.items()))
    print("   cycles/iter: " + " ".join(f"{k}={v/(n0 if k not in ('init','total') else 1):.0f}" for k, v in ph.items()), flush=True)

if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    ws = [synth.window_M(i, sigma_px=1.0, start="far") for i in range(8)]
    if which in ("all", "single"):
        for cs in (0, 64, 48, 32, 24, 16, 8):
            run(ws[:1], cs)
    if which in ("all", "batch"):
        for cs in (0, 18, 16, 12, 10, 8):
            run(ws, cs)
    if which in ("batch0",):
        run(ws, 0); run(ws, 0); run(ws[:1], 0)
    if which in ("scale",):
        # throughput against windows per launch (smaller groups, more windows resident)
        wl = [synth.window_M(i, sigma_px=1.0, start="far") for i in range(64)]
        for n, cs in ((8, 0), (16, 0), (16, 8), (32, 0), (64, 0), (64, 1)):
            run(wl[:n], cs, reps=5)
    if which in ("all", "small"):
        s = [synth.window_S(i, sigma_px=1.0, start="far") for i in range(8)]
        for cs in (0, 16, 8, 4, 1):
            run(s[:1], cs)
