#!/usr/bin/env python
"""Aggregate an `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv` dump per CUDA source line:
stall samples, executed warp instructions and the dominant stall reasons.  Usage: ncu_lines.py dump.csv [top]"""
import csv
import sys
from collections import defaultdict


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    agg = defaultdict(lambda: defaultdict(float))
    text = {}
    fname = ""
    hdr = None
    for r in rows:
        if len(r) == 2 and r[0] == "File Name":
            fname = r[1].split("/")[-1]
            continue
        if len(r) > 10 and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or r[0] == "":
            continue   # blank line number = the SASS rows under a CUDA line (already aggregated on the line's own row)
        key = (fname, r[0])
        text[key] = r[1].strip()
        for name, v in zip(hdr[4:], r[4:]):
            try:
                agg[key][name] += float(v)
            except ValueError:
                pass
    tot_s = sum(a["# Samples"] for a in agg.values()) or 1
    tot_i = sum(a["Instructions Executed"] for a in agg.values()) or 1
    print(f"total samples {tot_s:.0f}  total warp instructions {tot_i:.0f}")
    stall_names = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot_st = {s: sum(a[s] for a in agg.values()) for s in stall_names}
    print("stall mix:", ", ".join(f"{s[6:]} {100*v/tot_s:.1f}%" for s, v in sorted(tot_st.items(), key=lambda kv: -kv[1])[:8]))
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
        st = sorted(((a[s], s[6:]) for s in stall_names), reverse=True)[:2]
        print(f"{100*a['# Samples']/tot_s:5.1f}% smp {100*a['Instructions Executed']/tot_i:5.1f}% ins  {key[0]}:{key[1]:>4s}  "
              f"[{st[0][1]} {st[0][0]:.0f}, {st[1][1]} {st[1][0]:.0f}]  {text[key][:100]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
