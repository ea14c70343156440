#!/usr/bin/env python
"""Attributes the executed SASS instructions of an ncu report to CUDA source lines.
usage: ncu_by_line.py report.ncu-rep object.o kernel-substring [top]
Joins `ncu --page source --csv` (per-SASS-instruction executed counts, in address order) with
`nvdisasm --print-line-info` of the same object (per-instruction file:line)."""
import csv
import re
import subprocess
import sys
import collections


def main():
    rep, obj, kname = sys.argv[1:4]
    import os
    obj = os.path.abspath(obj)
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    import os as _os
    sel = ["-k", "regex:" + _os.environ["NCU_K"]] if _os.environ.get("NCU_K") else []  # pick one kernel of a multi-kernel report
    sel += ["--launch-count", "1"] if sel else []
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + sel, capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    iex, isrc, ist = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
    counts = [(r[isrc].strip(), int(r[iex]), int(r[ist])) for r in rows[2:] if len(r) > iex and r[iex].isdigit()]
    # line info from nvdisasm
    cub = subprocess.run(["cuobjdump", "-lelf", obj], capture_output=True, text=True).stdout
    elf = [l.split()[-1] for l in cub.splitlines() if "ELF file" in l][0]
    subprocess.run(["cuobjdump", "-xelf", elf, obj], cwd="/tmp", capture_output=True)
    dis = subprocess.run(["nvdisasm", "--print-line-info-inline", "/tmp/" + elf], capture_output=True, text=True).stdout
    lines = []
    infunc = False
    cur = "?"
    fresh, chain = True, []
    for l in dis.splitlines():
        if l.startswith(".text."):
            infunc = kname in l
            continue
        if not infunc:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            # one line per frame, innermost first: "File A, line n inlined at B, line m" / "File B, line m inlined at C ..."
            frame = f"{m.group(1).split('/')[-1]}:{m.group(2)}"
            if fresh:
                chain = [frame]
                fresh = False
            else:
                chain.append(frame)
            tail = re.search(r'inlined at "([^"]+)", line (\d+)', l)
            outer = f"{tail.group(1).split('/')[-1]}:{tail.group(2)}" if tail else None
            cur = " <- ".join(chain + ([outer] if outer else []))
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(.*?);", l)
        if m:
            lines.append((cur, m.group(2).strip()))
            fresh = True
    if len(lines) != len(counts):
        print(f"warning: {len(lines)} disassembled vs {len(counts)} profiled instructions", file=sys.stderr)
    agg = collections.Counter()
    stall = collections.Counter()
    for (where, _), (_, n, st) in zip(lines, counts):
        agg[where] += n
        stall[where] += st
    tot = sum(agg.values())
    tst = sum(stall.values()) or 1
    print(f"total warp instructions {tot}")
    # also aggregate by outermost inlined-at (the switch case line in ab_interp.cuh)
    outer = collections.Counter()
    for k, v in agg.items():
        outer[k.split(" <- ")[-1]] += v
    print("--- by call site in the interpreter (outermost frame) ---")
    for k, v in outer.most_common(top):
        print(f"{100 * v / tot:6.2f}%  {k}")
    if os.environ.get("NCU_BY_OPCODE"):  # executed instructions per (outermost line, opcode class)
        byop = collections.Counter()
        for (where, sass), (_, n, st) in zip(lines, counts):
            t = sass.split()
            op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
            byop[(where.split(" <- ")[-1], op)] += n
        print("--- by (outermost line, opcode) ---")
        for (w, op), v in byop.most_common(top):
            print(f"{100 * v / tot:6.2f}%  {w}  {op}")
    print("--- by innermost line ---")
    for k, v in agg.most_common(top):
        print(f"{100 * v / tot:6.2f}%  stall {100 * stall[k] / tst:5.1f}%  {k}")


if __name__ == "__main__":
    main()
