#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` (SASS view) dump by CUDA source line, using nvdisasm -g line info.

usage: ncu_by_line.py <report.ncu-rep> <lib.so> <kernel-substring> [top]
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def main():
    rep, so, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
    hdr = rows[h]
    ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall_cols = [(i, n) for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
    prof = {}
    base = None
    for r in rows[h + 1:]:
        if len(r) < len(hdr):
            continue
        try:
            a = int(r[ia], 16)
        except ValueError:
            continue
        if base is None:
            base = a
        prof[a - base] = (int(r[ii] or 0), int(r[isamp] or 0), {n: int(r[i] or 0) for i, n in stall_cols})
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
    line_of = {}
    for f in os.listdir(tmp):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        active, cur = False, ("?", 0)
        for ln in txt.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
            if m:
                active = kern in m.group(1)
                continue
            if not active:
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
            if m:
                line_of[int(m.group(1), 16)] = cur
    agg = defaultdict(lambda: [0, 0, defaultdict(int)])
    for off, (n, s, st) in prof.items():
        k = line_of.get(off, ("?", 0))
        agg[k][0] += n
        agg[k][1] += s
        for kk, v in st.items():
            agg[k][2][kk] += v
    tot_i = sum(v[0] for v in agg.values())
    tot_s = sum(v[1] for v in agg.values())
    print(f"total warp-instructions {tot_i}  samples {tot_s}  sass instrs {len(prof)}")
    src_cache = {}
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        st = sorted(v[2].items(), key=lambda x: -x[1])[:3]
        text = ""
        for root in ("slslam_b200/csrc",):
            p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", root, k[0])
            if os.path.exists(p):
                if p not in src_cache:
                    src_cache[p] = open(p).read().splitlines()
                if 0 < k[1] <= len(src_cache[p]):
                    text = src_cache[p][k[1] - 1].strip()[:90]
        print(f"{k[0]}:{k[1]:<5d} inst {100*v[0]/tot_i:5.1f}%  samp {100*v[1]/tot_s:5.1f}%  {' '.join(f'{a[6:]}={b}' for a,b in st):40s} | {text}")


if __name__ == "__main__":
    main()
