#!/usr/bin/env python
"""Warp-stall samples and executed instructions of one kernel by source line (ncu --import-source on, -lineinfo build).

    python tools/ncu_source_lines.py report.ncu-rep k_walk [top] > profiles/..._by_line.txt
"""
import csv
import subprocess
import sys
from collections import defaultdict


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern + "$"],
                         capture_output=True, text=True).stdout
    path = None
    hdr = None
    lines = {}
    per_file = defaultdict(lambda: [0, 0])
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            path = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or path is None or r[0] == "":
            continue
        d = dict(zip(hdr[4:], r[4:]))          # (two columns are called "Source": skip the first four)
        try:
            smp = int(d["# Samples"]); ins = int(d["Instructions Executed"])
        except (KeyError, ValueError):
            continue
        key = (path, int(r[0]))
        st = {k: int(v) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k and v.isdigit()}
        if key in lines:
            lines[key][0] += smp; lines[key][1] += ins
            for k, v in st.items():
                lines[key][3][k] = lines[key][3].get(k, 0) + v
        else:
            lines[key] = [smp, ins, r[1].strip(), st]
        per_file[path][0] += smp; per_file[path][1] += ins
    tot_s = sum(v[0] for v in lines.values()); tot_i = sum(v[1] for v in lines.values())
    print("kernel %s: %d samples, %d warp instructions" % (kern, tot_s, tot_i))
    for p, (s, i) in sorted(per_file.items(), key=lambda x: -x[1][0]):
        print("  %-22s samples %5.1f%%  inst %5.1f%%" % (p, 100.0 * s / max(tot_s, 1), 100.0 * i / max(tot_i, 1)))
    print("--- top lines by samples")
    for (p, ln), (s, i, src, st) in sorted(lines.items(), key=lambda x: -x[1][0])[:top]:
        main_st = sorted(st.items(), key=lambda x: -x[1])[:2]
        print("%5.1f%% smp %5.1f%% inst  %s:%d  [%s]  %s" % (100.0 * s / max(tot_s, 1), 100.0 * i / max(tot_i, 1), p, ln,
                                                           " ".join("%s %d" % (k[6:], v) for k, v in main_st), src[:110]))


if __name__ == "__main__":
    main()
