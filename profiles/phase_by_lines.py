#!/usr/bin/env python
"""Sum an ncu source page (ncu -i X.ncu-rep --page source --csv) over the phases of k_score_fused: every SASS
instruction is attributed to the source line nvdisasm gives it; helper lines inlined from headers inherit the phase
of the last pk_fused.cu line seen before them in address order. Phases are ranges of pk_fused.cu lines:
usage: phase_by_lines.py <source.csv> <nvdisasm --print-line-info dump> <mangled kernel name> "<line>:<phase>,<line>:<phase>,..."
(find the line numbers with:  grep -n "T1 (TM\\|A3a (TM\\|A2 (TM\\|A3b (TM\\|---- A1:\\|---- A2:\\|---- A3:\\|---- A4:\\|---- A5:\\|phase B" pk_fused.cu)"""
import collections
import csv
import re
import sys

sass_csv, dis, kname, spec = sys.argv[1:5]
bounds = [(0, "setup")] + [(int(a), b) for a, b in (x.split(":") for x in spec.split(","))]
addr2line, cur, infn = {}, None, False
for ln in open(dis):
    if ln.startswith(".text."):
        infn = ln.strip().rstrip(":") == ".text." + kname
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
    if m and cur:
        addr2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(sass_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci = {k: hdr.index(k) for k in hdr}
KEYS = ("# Samples", "Instructions Executed", "L1 Wavefronts Shared", "stall_long_sb", "stall_short_sb", "stall_barrier",
        "stall_wait", "stall_math", "stall_mio", "stall_not_selected", "stall_selected")


def phase_of(line):
    p = bounds[0][1]
    for b, n in bounds:
        if line >= b:
            p = n
    return p


agg, base, last = collections.defaultdict(collections.Counter), None, "setup"
for r in rows[hi + 1:]:
    if r and r[0] == "Address":
        break
    if len(r) < len(hdr):
        continue
    a = int(r[ci["Address"]], 16)
    base = a if base is None else base
    f, l = addr2line.get(a - base, ("?", 0))
    if f == "pk_fused.cu" and l >= bounds[1][0]:
        last = phase_of(l)
    for k in KEYS:
        try:
            agg[last][k] += int(float(r[ci[k]]))
        except ValueError:
            pass
tot = collections.Counter()
for c in agg.values():
    tot.update(c)
print("| phase | warp-inst | % | samples | % | smem wavefronts | long_sb | short_sb | barrier | wait | math | mio | not_selected | selected |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
seen = []
for b, n in bounds:
    if n in seen or n not in agg:
        continue
    seen.append(n)
    c = agg[n]
    print("| %s | %d | %.1f | %d | %.1f | %d | %d | %d | %d | %d | %d | %d | %d | %d |" % (
        n, c["Instructions Executed"], 100.0 * c["Instructions Executed"] / tot["Instructions Executed"], c["# Samples"],
        100.0 * c["# Samples"] / tot["# Samples"], c["L1 Wavefronts Shared"], c["stall_long_sb"], c["stall_short_sb"],
        c["stall_barrier"], c["stall_wait"], c["stall_math"], c["stall_mio"], c["stall_not_selected"], c["stall_selected"]))
