#!/usr/bin/env python
"""Motion-only BA (reference src/slam.cpp:578-675: one free camera, every line constant), the per-frame call:
end-to-end latency through the C ABI -- dedicated kernel (default) and general kernel (SLSLAM_NO_MOBA_FASTPATH=1) --
against the single-thread oracle."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from slslam_b200 import capi, synth
from oracle import oracle


def timed(w, reps=50):
    for _ in range(5):
        capi.lba_solve(w, max_iters=10)
    ts, ks = [], []
    for _ in range(reps):
        t0 = time.perf_counter(); p, s = capi.lba_solve(w, max_iters=10); ts.append(time.perf_counter() - t0)
        ks.append(capi.last_timings()["dev_kernel_ms"])
    return np.median(ts) * 1e3, np.median(ks), s


for nl in (60, 200, 400):
    w = synth.motion_only_window(11, num_lines=nl)
    tf, kf, sf = timed(w)
    os.environ["SLSLAM_NO_MOBA_FASTPATH"] = "1"
    tg, kg, sg = timed(w)
    del os.environ["SLSLAM_NO_MOBA_FASTPATH"]
    tc = []
    for _ in range(20):
        t0 = time.perf_counter(); po, so = oracle.lba_solve(w, max_iters=10, solver=1); tc.append(time.perf_counter() - t0)
    print(f"motion-only BA, {w.num_lines} lines / {w.num_observations} obs, {sf['iterations']} LM iterations: "
          f"dedicated kernel e2e {tf:.3f} ms (kernel {kf:.3f} ms) | general kernel e2e {tg:.3f} ms (kernel {kg:.3f} ms) | "
          f"oracle 1 thread {np.median(tc)*1e3:.3f} ms | final cost rel diff {abs(sf['final_cost']-so['final_cost'])/so['final_cost']:.1e}",
          flush=True)
# a batch of frames in one launch (one CTA per frame)
ws = [synth.motion_only_window(100 + i, num_lines=200) for i in range(64)]
for _ in range(3):
    capi.lba_solve_batch(ws, max_iters=10)
t0 = time.perf_counter()
for _ in range(10):
    ps, ss = capi.lba_solve_batch(ws, max_iters=10)
dt = (time.perf_counter() - t0) / 10
print(f"batch of 64 motion-only problems (200 lines each): {dt*1e3:.3f} ms per batch e2e, kernel {capi.last_timings()['dev_kernel_ms']:.3f} ms, "
      f"{sum(s['iterations'] for s in ss)/dt:.0f} LM iterations/s")
