#!/usr/bin/env python
"""Aggregate ncu source-page samples of k_score_fused by kernel phase.
usage: phase_samples.py <sass.csv> <nvdisasm --print-line-info dump> <mangled kernel> <pk_fused.cu>
Phases are found from the marker comments in pk_fused.cu; inlined helpers are attributed by
their call-site phase when nvdisasm gives an inline chain, else by helper name."""
import collections, csv, re, sys
sass_csv, dis, kname, src = sys.argv[1:5]
lines = open(src).read().split("\n")
marks = []
for i, l in enumerate(lines, 1):
    m = re.search(r"// ---- (A\d)[:a-z ]|// =+ phase (B)|// ---- one-time setup", l)
    if m:
        marks.append((i, m.group(1) or m.group(2) or "setup"))
def phase_of(line):
    p = "helpers"
    for ln, name in marks:
        if line >= ln:
            p = name
    return p if line >= marks[0][0] else "helpers"
addr2line, cur, infn = {}, None, False
for ln in open(dis):
    if ln.startswith(".text."):
        infn = ln.strip().rstrip(":") == ".text." + kname
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        f, l = m.group(1).split("/")[-1], int(m.group(2))
        if m.group(3) and m.group(3).endswith("pk_fused.cu"):
            cur = ("pk_fused.cu", int(m.group(4)), f, l)
        else:
            cur = (f, l, f, l)
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
    if m and cur:
        addr2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(sass_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
cols = ["# Samples", "Instructions Executed", "L1 Wavefronts Shared", "stall_wait", "stall_short_sb", "stall_long_sb",
        "stall_barrier", "stall_math", "stall_mio", "stall_not_selected", "stall_selected", "stall_lg", "stall_branch_resolving", "stall_no_inst"]
ci = {k: hdr.index(k) for k in cols + ["Address"]}
agg = collections.defaultdict(collections.Counter)
unk = collections.Counter()
base = None
curph = "setup"
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    a = int(r[ci["Address"]], 16)
    if base is None:
        base = a
    loc = addr2line.get(a - base)
    # helper / library lines inherit the phase of the surrounding kernel-body code (SASS is in program order)
    if loc is not None and loc[0] == "pk_fused.cu" and phase_of(loc[1]) != "helpers":
        curph = phase_of(loc[1])
    ph = curph
    for k in cols:
        v = r[ci[k]]
        agg[ph][k] += int(float(v)) if v else 0
tot = sum(v["# Samples"] for v in agg.values())
print("%-8s %8s %6s %10s %10s | %s" % ("phase", "samples", "%", "warp-inst", "smem-wf", " ".join(c.replace("stall_", "")[:8].rjust(8) for c in cols[3:])))
for ph, v in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"]):
    print("%-8s %8d %6.1f %10d %10d | %s" % (ph, v["# Samples"], 100.0 * v["# Samples"] / tot, v["Instructions Executed"],
                                              v["L1 Wavefronts Shared"], " ".join(str(v[c]).rjust(8) for c in cols[3:])))
if unk:
    print("unattributed helper lines:", unk.most_common(8))
