#!/usr/bin/env python
"""Walk the SASS source page of an ncu report: stall-sample totals per reason and the
instructions that collect the samples, in address order with landmark instructions.

    ncu -i rep.ncu-rep --page source --csv > src.csv ; python tools/ncu_walk.py src.csv [kernel-index] [min-samples]
"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
thr = int(sys.argv[3]) if len(sys.argv) > 3 else 15
starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
names = [rows[i - 1][1] if i > 0 else "?" for i in starts]
for n, nm in enumerate(names):
    print(f"[{n}] {nm}")
h = rows[starts[which]]
end = starts[which + 1] - 1 if which + 1 < len(starts) else len(rows)
ci = {n: i for i, n in enumerate(h)}
data = [r for r in rows[starts[which] + 1:end] if len(r) >= len(h)]
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]


def samples(r):
    try:
        return int(r[ci["# Samples"]])
    except ValueError:
        return 0


tot = collections.Counter()
for r in data:
    for s in stalls:
        try:
            tot[s] += int(r[ci[s]])
        except ValueError:
            pass
print("total samples", sum(samples(r) for r in data), tot.most_common(8))
marks = ("UTCHMMA", "BAR.", "UBLKCP", "STTM", "UTCBAR", "LDTM")
seg = 0
prev_mark = None
for k, r in enumerate(data):
    src = r[ci["Source"]].strip()
    s = samples(r)
    seg += s
    m = [x for x in marks if x in src]
    if s >= thr or (m and m[0] != prev_mark):
        print(f"{k:5d} seg={seg:5d} s={s:4d} ex={r[ci['Instructions Executed']]:>7} {src[:84]}")
        seg = 0
    if m:
        prev_mark = m[0]
