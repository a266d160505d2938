#!/usr/bin/env python
"""Instruction / stall-sample share per code region of pb_scl_lut.cuh from an ncu cuda,sass source CSV.
Regions are found by marker comments in the kernel source, so the script follows code motion.
Usage: ncu_regions.py dump.csv kernel_source.cuh passes"""
import csv
import sys
from collections import defaultdict

MARKERS = [
    ("ring / next_line", "// ---- table stream."),
    ("frame setup + in8", "const long long n_groups"),
    ("ptr helpers", "double PM = (me == 0)"),
    ("fg_step (upper f/g)", "// f / g step at depth dd"),
    ("combine (upper)", "// combine at depth dc"),
    ("subtree control + A/B lookups", "// ---- by-value state of the 8-leaf subtree"),
    ("leaf symbol + LLR", "// leaf symbol through the depth n-1 node"),
    ("leaf decision / fork", "// leaf decision (PD/src"),
    ("subtree combines + publish", "xb = (xb & ~(1u << lp))"),
    ("epilogue", "// ---------------- epilogue"),
    ("host", "// host side"),
]


def main(path, srcpath, passes):
    src = open(srcpath).read().split("\n")
    bounds = []
    for name, mark in MARKERS:
        for i, l in enumerate(src):
            if mark in l:
                bounds.append((i + 1, name))
                break
    bounds.sort()

    def region(ln, text):
        if ln - 1 >= len(src) or src[ln - 1].strip() != text.strip():
            return "inlined intrinsics (shfl, sync, cvta)"
        name = "prologue"
        for b, nm in bounds:
            if ln >= b:
                name = nm
        return name

    rows = list(csv.reader(open(path)))
    hdr = None
    ins, smp = defaultdict(float), defaultdict(float)
    for r in rows:
        if len(r) > 10 and r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or r[0] == "":
            continue
        try:
            i, s, ln = float(r[7]), float(r[6]), int(r[0])
        except ValueError:
            continue
        k = region(ln, r[1])
        ins[k] += i
        smp[k] += s
    ti, ts = sum(ins.values()), sum(smp.values())
    print(f"total warp instructions {ti:.0f} = {ti/passes:.0f} per pass ({passes} passes)")
    for k, v in sorted(ins.items(), key=lambda kv: -kv[1]):
        print(f"{k:40s} {100*v/ti:5.1f}% ins {100*smp[k]/ts:5.1f}% smp {v/passes:9.0f} instr/pass")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], float(sys.argv[3]))
