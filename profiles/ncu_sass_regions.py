#!/usr/bin/env python
"""Instruction / stall-sample share per code region of pb_scl_lut.cuh from an
`ncu -i X.ncu-rep --page source --print-source cuda,sass --csv` dump, attributed per SASS instruction: instructions that
belong to inlined intrinsics (shuffles, syncs: other source files) take the region of the nearest preceding instruction of
the kernel source in address order.  Usage: ncu_sass_regions.py dump.csv kernel_source.cuh frames"""
import csv
import sys
from collections import defaultdict

MARKERS = [
    ("ring / next_line", "// ---- table stream."),
    ("frame setup + in8", "const long long n_groups"),
    ("ptr helpers", "double PM = (me == 0)"),
    ("fg_step (upper f/g)", "// f / g step at depth dd"),
    ("combine (upper)", "// combine at depth dc"),
    ("fork", "auto fork = [&]"),
    ("special nodes (Fast)", "auto special = [&]"),
    ("op loop", "for (int oi = 0; oi < fp.n_ops; ++oi)"),
    ("SUB8 subtree", "case FOP_SUB8"),
    ("micro-op subtrees", "case FOP_SBEGIN"),
    ("micro-ops (after SUB8)", "case FOP_SF3"),
    ("epilogue", "// ---------------- epilogue"),
    ("host", "// host side"),
]


def main(path, srcpath, frames):
    src = open(srcpath).read().split("\n")
    bounds = []
    for name, mark in MARKERS:
        for i, l in enumerate(src):
            if mark in l:
                bounds.append((i + 1, name))
                break
    bounds.sort()

    def region(ln):
        name = "prologue / stream accessors"
        for b, nm in bounds:
            if ln >= b:
                name = nm
        return name

    rows = list(csv.reader(open(path)))
    sass = []       # (address, region or None, samples, instructions, opcode)
    cur = None
    for r in rows:
        if len(r) > 3 and r[0] not in ("", "Line No"):
            try:
                ln = int(r[0])
            except ValueError:
                continue
            txt = r[1].strip()
            own = ln - 1 < len(src) and src[ln - 1].strip()[:30] == txt[:30] and txt != ""
            cur = region(ln) if own else None
            continue
        if len(r) >= 8 and r[0] == "" and r[2].startswith("0x"):
            try:
                sass.append((int(r[2], 16), cur, float(r[6]), float(r[7]), r[3].split()[0 if not r[3].strip().startswith("@") else 1]))
            except ValueError:
                pass
    uniq = {}       # an instruction of an inlined function is listed under the callee's line and under the call site
    for t in sass:
        if t[0] not in uniq or (uniq[t[0]][1] is None and t[1] is not None):
            uniq[t[0]] = t
    sass = sorted(uniq.values(), key=lambda t: t[0])
    ins, smp = defaultdict(float), defaultdict(float)
    ops = defaultdict(float)
    last = "prologue / stream accessors"
    for a, reg, s, i, op in sass:
        if reg is not None:
            last = reg
        ins[last] += i
        smp[last] += s
        ops[op.split(".")[0]] += i
    ti, ts = sum(ins.values()), sum(smp.values())
    print(f"total warp instructions {ti:.0f} = {ti/frames:.0f} per frame ({frames:.0f} frames)")
    for k, v in sorted(ins.items(), key=lambda kv: -kv[1]):
        print(f"{k:40s} {100*v/ti:5.1f}% ins {100*smp[k]/ts:5.1f}% smp {v/frames:9.0f} instr/frame")
    print("opcode mix:", ", ".join(f"{k} {100*v/ti:.1f}%" for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:22]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], float(sys.argv[3]))
