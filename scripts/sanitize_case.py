#!/usr/bin/env python
"""Small invocations of every kernel, for compute-sanitizer (memcheck / racecheck / synccheck) under gpurun."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
from slslam_b200 import capi, synth

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "lba"):
    w = synth.make_window(0, 6, 60, 260, sigma_px=0.5)
    p, s = capi.lba_solve(w, max_iters=3)                       # plan kernel + gather + solve kernel (group of CTAs)
    print("lba", s["iterations"], s["final_cost"])
    ws = [synth.make_window(i, 5, 40, 150) for i in range(3)]
    ps, ss = capi.lba_solve_batch(ws, max_iters=2)
    print("lba batch", [x["final_cost"] for x in ss])
if which in ("all", "moba"):
    p, s = capi.lba_solve(synth.motion_only_window(1, num_lines=40), max_iters=4)
    print("moba", s["iterations"], s["final_cost"])
if which in ("all", "ransac"):
    from test_ransac import make_case
    poses, lines, obs, _ = make_case(0, 70, 9)
    print("ransac", capi.ransac_score(poses, lines, obs)[0])
if which in ("all", "po"):
    g = synth.make_pose_graph(0, num_poses=16, neighbours=2, num_loops=2)
    p, s = capi.po_solve(g, max_iters=3)
    print("po", s["final_cost"])
if which in ("all", "po"):
    os.environ["SLSLAM_PO_COLUMNS"] = "1"                        # minimum-degree order, column-at-a-time kernel
    p, s = capi.po_solve(synth.make_pose_graph(0, num_poses=16, neighbours=2, num_loops=2), max_iters=2)
    del os.environ["SLSLAM_PO_COLUMNS"]
    print("po columns", s["final_cost"])
if which in ("all", "po"):
    os.environ["SLSLAM_PO_DENSE"] = "1"                          # the dense fallback path as well
    p, s = capi.po_solve(synth.make_pose_graph(0, num_poses=16, neighbours=2, num_loops=2), max_iters=2)
    del os.environ["SLSLAM_PO_DENSE"]
    print("po dense", s["final_cost"])
if which in ("all", "wide"):
    w = synth.make_window(31, 18, 60, 700, num_fixed_cameras=16, sigma_px=0.5)      # 34 camera blocks: the general kernel
    p, s = capi.lba_solve(w, max_iters=2)
    print("lba wide", w.num_cameras, s["iterations"], s["final_cost"])
if which in ("all", "map"):
    from all_kernels_driver import map_case
    map_case()
    print("map ok")
if which in ("all", "device"):
    import torch
    from slslam_b200 import shard
    ws = [synth.make_window(40 + i, 5, 40, 150) for i in range(2)]
    buf, lay = shard.pack_rank_buffer(ws, pin=True)
    dev = torch.from_numpy(buf).cuda()
    ss = capi.lba_solve_batch_device(lay.shapes, dev.data_ptr(), lay.offsets(), max_iters=2, summaries_dev_ptr=dev.data_ptr() + lay.summary_off,
                                     stream=torch.cuda.current_stream().cuda_stream)
    print("lba device", [x["final_cost"] for x in ss])
