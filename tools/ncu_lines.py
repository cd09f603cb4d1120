#!/usr/bin/env python
"""Warp-state samples of an ncu source page (SASS view) folded onto SOURCE LINES: the SASS order of
the report is aligned with `nvdisasm -g -c` of the same cubin (line markers), samples are summed
over all captured launches of the kernel.

    cuobjdump -xelf all libpgw_b200.so ; nvdisasm -g -c step_fused.sm_100a.cubin > sass.txt
    ncu -i rep.ncu-rep --page source --csv > src.csv
    python tools/ncu_lines.py src.csv sass.txt '_ZN3pgw17step_fused_kernelILb0EEEvNS_11FusedParamsE' [file-filter]
"""
import collections
import csv
import re
import sys

src_csv, sass_txt, mangled = sys.argv[1:4]
flt = sys.argv[4] if len(sys.argv) > 4 else None

# nvdisasm: instruction -> (file, line), inlined-at chain ignored (innermost position)
lines = open(sass_txt).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text." + mangled + ":"))
pos, cur = [], ("?", 0)
for l in lines[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"):
        break
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
        pos.append((cur, l.split("*/", 1)[1].strip()))

rows = list(csv.reader(open(src_csv)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
per_line = collections.Counter()
per_line_reason = collections.defaultdict(collections.Counter)
total = 0
launches = 0
for si, s in enumerate(starts):
    h = rows[s]
    ci = {n: i for i, n in enumerate(h)}
    end = starts[si + 1] - 1 if si + 1 < len(starts) else len(rows)
    data = [r for r in rows[s + 1:end] if len(r) >= len(h)]
    if len(data) != len(pos):
        continue
    launches += 1
    stalls = [n for n in h if n.startswith("stall_")]
    for (fl, _), r in zip(pos, data):
        try:
            n = int(r[ci["# Samples"]])
        except ValueError:
            n = 0
        per_line[fl] += n
        total += n
        for st in stalls:
            try:
                per_line_reason[fl][st] += int(r[ci[st]])
            except ValueError:
                pass
print(f"{launches} launches, {len(pos)} SASS instructions, {total} samples")
acc = 0
for (f, ln), n in sorted(per_line.items(), key=lambda kv: (kv[0][0], kv[0][1])):
    if flt and flt not in f:
        continue
    if n == 0:
        continue
    top = ", ".join(f"{k[6:]} {v}" for k, v in per_line_reason[(f, ln)].most_common(3) if v)
    print(f"{f}:{ln:<5d} {n:6d} {100.0 * n / total:5.1f}%   {top}")

# ---- the same samples in SASS (execution) order: runs of instructions of one source line
if "--order" in sys.argv:
    thr = int(sys.argv[sys.argv.index("--order") + 1])
    h = rows[starts[0]]
    ci = {n: i for i, n in enumerate(h)}
    sums = [0] * len(pos)
    reasons = [collections.Counter() for _ in pos]
    for si, s in enumerate(starts):
        end = starts[si + 1] - 1 if si + 1 < len(starts) else len(rows)
        data = [r for r in rows[s + 1:end] if len(r) >= len(h)]
        if len(data) != len(pos):
            continue
        for k, r in enumerate(data):
            try:
                sums[k] += int(r[ci["# Samples"]])
            except ValueError:
                pass
            for st in [n for n in h if n.startswith("stall_") and "Not Issued" not in n]:
                try:
                    reasons[k][st] += int(r[ci[st]])
                except ValueError:
                    pass
    print("---- SASS order")
    run_start, run_line, run_n, run_r, hot = 0, pos[0][0], 0, collections.Counter(), (0, "")
    cum = 0
    for k, ((fl, ins), n) in enumerate(zip(pos, sums)):
        if fl != run_line:
            if run_n >= thr:
                top = ", ".join(f"{a[6:]} {b}" for a, b in run_r.most_common(2))
                print(f"{run_start:5d}-{k - 1:5d} {run_line[0]}:{run_line[1]:<4d} {run_n:6d} ({100.0 * cum / total:5.1f}% cum)  {top}   | {hot[1][:60]}")
            run_start, run_line, run_n, run_r, hot = k, fl, 0, collections.Counter(), (0, "")
        run_n += n
        cum += n
        run_r.update(reasons[k])
        if n > hot[0]:
            hot = (n, ins)
