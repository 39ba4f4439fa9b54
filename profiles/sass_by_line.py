#!/usr/bin/env python
"""Join an ncu SASS page (ncu -i X.ncu-rep --page source --csv) with nvdisasm line
info to get executed warp-instructions and stall samples per CUDA source line.
usage: sass_by_line.py <sass.csv> <nvdisasm --print-line-info dump> <mangled kernel name> [top]"""
import collections
import csv
import re
import sys

sass_csv, dis, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
sort_key = sys.argv[5] if len(sys.argv) > 5 else "Instructions Executed"   # or "# Samples"
# address -> line from nvdisasm
addr2line = {}
cur = None
infn = False
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
ci = {k: hdr.index(k) for k in ("Address", "Source", "# Samples", "Instructions Executed", "Thread Instructions Executed",
                                 "L1 Wavefronts Shared", "stall_wait", "stall_long_sb", "stall_short_sb", "stall_barrier")}
base = None
agg = collections.defaultdict(lambda: collections.Counter())
tot = collections.Counter()
for r in rows[hi + 1:]:
    if r and r[0] == "Address":          # a second launch of the same kernel in the report: the first one is enough
        break
    if len(r) < len(hdr):
        continue
    a = int(r[ci["Address"]], 16)
    if base is None:
        base = a
    key = addr2line.get(a - base, ("?", 0))
    for k in ("# Samples", "Instructions Executed", "Thread Instructions Executed", "L1 Wavefronts Shared", "stall_wait",
              "stall_long_sb", "stall_short_sb", "stall_barrier"):
        try:
            v = int(float(r[ci[k]]))
        except ValueError:
            v = 0
        agg[key][k] += v
        tot[k] += v
print("total warp-instructions %d, samples %d, smem wavefronts %d" % (tot["Instructions Executed"], tot["# Samples"], tot["L1 Wavefronts Shared"]))
print("%-22s %12s %6s %8s %10s %8s %8s %8s %8s" % ("line", "warp-inst", "%", "samples", "smem-wf", "wait", "long_sb", "short_sb", "barrier"))
for key, c in sorted(agg.items(), key=lambda kv: -kv[1][sort_key])[:top]:
    print("%-22s %12d %6.2f %8d %10d %8d %8d %8d %8d" % ("%s:%d" % key, c["Instructions Executed"],
          100.0 * c["Instructions Executed"] / max(tot["Instructions Executed"], 1), c["# Samples"], c["L1 Wavefronts Shared"],
          c["stall_wait"], c["stall_long_sb"], c["stall_short_sb"], c["stall_barrier"]))
