"""Per-source-line / per-region breakdown of an ncu report (test/profiling infrastructure).

  ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > X.csv
  python tools/ncu_breakdown.py X.csv [file-substring marker=text ...]

Prints, per file and per region of the main file (regions = lines between marker comments "// ---- A." etc.),
warp instructions, active lanes, share of stall samples and the top stall reasons; then the hottest lines."""
from __future__ import annotations

import csv
import re
import sys
from collections import defaultdict


def load(path):
    rows = []
    cur_file, hdr = None, None
    for r in csv.reader(open(path, newline="")):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or r[0] == "" or not r[0].isdigit():
            continue
        d = dict(zip(hdr[4:], r[4:]))   # metrics (the two "Source" columns collide: skip the first four)
        rows.append((cur_file, int(r[0]), r[1], d))
    return rows


def num(d, k):
    try:
        return float(d.get(k, "0") or 0)
    except ValueError:
        return 0.0


STALLS = ["stall_barrier", "stall_branch_resolving", "stall_long_sb", "stall_short_sb", "stall_wait", "stall_selected",
          "stall_not_selected", "stall_math", "stall_mio", "stall_no_inst", "stall_lg", "stall_sleep", "stall_membar",
          "stall_dispatch"]


def summarize(name, items, tot_inst, tot_smp):
    inst = sum(num(d, "Instructions Executed") for d in items)
    thr = sum(num(d, "Thread Instructions Executed") for d in items)
    smp = sum(num(d, "# Samples") for d in items)
    st = {s: sum(num(d, s) for d in items) for s in STALLS}
    top = sorted(st.items(), key=lambda kv: -kv[1])[:4]
    tops = " ".join(f"{k[6:]}:{100 * v / max(smp, 1):.0f}%" for k, v in top if v)
    print(f"{name:28s} inst {inst / 1e9:7.2f}G ({100 * inst / max(tot_inst, 1):5.1f}%) lanes {thr / max(inst, 1):5.1f} "
          f"samples {100 * smp / max(tot_smp, 1):5.1f}%  {tops}")


def main():
    path = sys.argv[1]
    main_file = sys.argv[2] if len(sys.argv) > 2 else "inflate_cells.cuh"
    rows = load(path)
    tot_inst = sum(num(d, "Instructions Executed") for _, _, _, d in rows)
    tot_smp = sum(num(d, "# Samples") for _, _, _, d in rows)
    print(f"total {tot_inst / 1e9:.3f} G warp instructions, {int(tot_smp)} samples")
    by_file = defaultdict(list)
    for f, ln, src, d in rows:
        by_file[f.split("/")[-1]].append(d)
    for f, items in sorted(by_file.items(), key=lambda kv: -sum(num(d, "# Samples") for d in kv[1])):
        summarize("F:" + f, items, tot_inst, tot_smp)
    # regions of the main file: a new region starts at every line holding "// ----" or at a function definition
    lines = sorted((ln, src, d) for f, ln, src, d in rows if f.endswith(main_file))
    full = [f for f, _, _, _ in rows if f.endswith(main_file)]
    marks = []   # (line, name): marker comments "// ---- X" and function heads, read from the source file itself
    try:
        for i, text in enumerate(open(full[0]).read().split("\n"), 1):
            m = re.search(r"// ---- (.{1,24})", text)
            f = re.match(r"^(?:__device__|template|inline|__global__|struct)\b.*?(\w+)\s*[({]", text)
            if m:
                marks.append((i, m.group(1).strip()))
            elif f and not text.startswith(" "):
                marks.append((i, "fn " + f.group(1)))
    except (OSError, IndexError):
        pass
    regions, cur, mi = [], ("prologue", []), 0
    for ln, src, d in lines:
        while mi < len(marks) and marks[mi][0] <= ln:
            regions.append(cur)
            cur = (f"{marks[mi][0]}: {marks[mi][1][:22]}", [])
            mi += 1
        cur[1].append(d)
    regions.append(cur)
    print(f"--- regions of {main_file}")
    for name, items in regions:
        if items:
            summarize(name, items, tot_inst, tot_smp)
    print("--- hottest lines (by samples)")
    for f, ln, src, d in sorted(rows, key=lambda r: -num(r[3], "# Samples"))[:45]:
        smp, inst, thr = num(d, "# Samples"), num(d, "Instructions Executed"), num(d, "Thread Instructions Executed")
        st = sorted(((s, num(d, s)) for s in STALLS), key=lambda kv: -kv[1])[:2]
        tops = " ".join(f"{k[6:]}:{100 * v / max(smp, 1):.0f}%" for k, v in st if v)
        print(f"{f.split('/')[-1][:16]:16s}:{ln:5d} {inst / 1e9:6.2f}G lanes {thr / max(inst, 1):5.1f} smp {100 * smp / max(tot_smp, 1):5.1f}% {tops:32s} {src.strip()[:90]}")


if __name__ == "__main__":
    main()
