#!/usr/bin/env python
"""Collects the reference's published house-simulation LBA results (the only numbers the reference ships for this path)
into tests/golden/house_ba_results.json: for every /root/reference/matlab_script/result_comp_ancdir_orthonorm/
ba_result_<param>_err<sigma>_basize<W>_maxnumiter<M>.txt the four lines written by src/main.cpp:84-89
(average iterations per frame, total time, average initial cost, average final cost).  Run in the build container;
the fixture travels, /root/reference does not."""
import glob, json, os, re
out = []
for f in sorted(glob.glob("/root/reference/matlab_script/result_comp_ancdir_orthonorm/ba_result_*.txt")):
    m = re.match(r"ba_result_(\w+)_err([\d.]+)_basize(\d+)_maxnumiter(\d+)\.txt", os.path.basename(f))
    v = [float(l.split("=")[1]) for l in open(f).read().strip().splitlines()]
    out.append(dict(parameterisation=m.group(1), sigma_px=float(m.group(2)), window=int(m.group(3)), max_iter=int(m.group(4)),
                    mean_iterations=v[0], total_time_s=v[1], mean_initial_cost=v[2], mean_final_cost=v[3]))
json.dump(out, open(os.path.join(os.path.dirname(__file__), "house_ba_results.json"), "w"), indent=0)
print(len(out), "records")
