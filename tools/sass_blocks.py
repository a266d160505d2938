#!/usr/bin/env python
"""Static view of one kernel's SASS: basic blocks (split at labels and branches) with their instruction count and the
CUDA source lines they come from.  Used with known trip counts to budget instructions per decoded frame without a GPU.
Usage: sass_blocks.py file.cubin kernel-name-substring [min_instrs]"""
import re
import subprocess
import sys
from collections import Counter


def main():
    cubin, pat = sys.argv[1], sys.argv[2]
    minins = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    txt = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.split("\n")
    on = False
    blocks, cur, line = [], None, 0
    for l in txt:
        if l.startswith(".text."):
            on = pat in l
            continue
        if not on:
            continue
        if l.startswith("//-----") and ".text." in l:
            on = False
            continue
        m = re.match(r'\s*//## File ".*?([^/"]+)", line (\d+)', l)
        if m:
            line = (m.group(1), int(m.group(2)))
            continue
        if re.match(r"^\.L_x_\d+:", l):
            cur = {"label": l.strip(), "ins": [], "lines": Counter()}
            blocks.append(cur)
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", l)
        if m:
            if cur is None:
                cur = {"label": "entry", "ins": [], "lines": Counter()}
                blocks.append(cur)
            ins = m.group(2).strip()
            cur["ins"].append(ins)
            cur["lines"][line] += 1
            op = re.sub(r"^@!?U?P\d+\s+", "", ins).split()[0]
            if op.startswith(("BRA", "EXIT", "RET", "BRX", "CALL")) and not ins.startswith("@"):
                cur = None
            elif op.startswith("BRA"):   # conditional branch ends the block too
                nxt = {"label": "(fallthrough)", "ins": [], "lines": Counter()}
                blocks.append(nxt)
                cur = nxt
    total = 0
    for b in blocks:
        n = len(b["ins"])
        total += n
        if n >= minins:
            ops = Counter(re.sub(r"^@!?U?P\d+\s+", "", i).split()[0].split(".")[0] for i in b["ins"])
            top = ", ".join(f"{f}:{ln}x{c}" for (f, ln), c in b["lines"].most_common(4))
            print(f"{b['label']:16s} {n:4d}  shfl={ops['SHFL']:2d} dsetp={ops['DSETP']:2d} lds={ops['LDS']:2d}  [{top}]")
    print("total static instructions:", total)


if __name__ == "__main__":
    main()
