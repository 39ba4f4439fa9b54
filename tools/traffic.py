#!/usr/bin/env python
"""profiles/roofline_traffic.json from ncu launch lists that carry DRAM counters:
   ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
       --log-file traffic_c2.csv python bench.py --workload c2 --steps 3 --warmup 3 --no-cpu-baseline --genome none
   python tools/traffic.py c2=traffic_c2.csv c4=traffic_c4.csv > profiles/roofline_traffic.json
Per workload and bench stage: median DRAM bytes (read + write) per launch of every kernel of the stage, summed."""
import collections
import csv
import json
import sys

import numpy as np

STAGE = [("k_score_fused", "features"), ("k_band_csr<int, int", "band_build"), ("k_band_rowmajor", "band_build"),
         ("k_band_csr<unsigned short", "band_build_uint16_columns"), ("k_band_rows", "band_build_packed_rows"),
         ("k_band_escapes", "band_build_packed_rows"), ("k_scatter_pixels", "band_build_coo"),
         ("k_valid_bits", "diag_sums"), ("k_diag_", "diag_sums"),
         ("k_fit_expected", "expected_fit"), ("k_cand_", "candidate_scan"), ("k_emit", "emit"), ("k_row_offsets", "emit"),
         ("k_record_", "emit"), ("k_features", "features_unfused"), ("k_forest", "forest_unfused")]
# band_build is the device-resident upload bench.py times per stage (cooler's int32 columns, plus the row-major copy
# where the fused kernel wants it); the other upload encodings of the end-to-end passes are listed beside it


def unit_scale(u):
    return {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "ns": 1e-3}.get(u, 1)


def main():
    out = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum per launch (median over the launches of a bench run under "
                       "ncu --clock-control none), per workload and bench stage; kernels lists the kernels of the stage. "
                       "Made by tools/traffic.py from the ncu launch lists named in `source`."}
    for arg in sys.argv[1:]:
        wl, path = arg.split("=", 1)
        rows = [r for r in csv.reader(open(path)) if len(r) > 10]
        h = rows[0]
        ki, mi, ui, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Unit"), h.index("Metric Value"), h.index("ID")
        per = collections.defaultdict(lambda: collections.defaultdict(dict))       # kernel -> launch id -> metric -> value
        for r in rows[1:]:
            try:
                v = float(r[vi].replace(",", "")) * unit_scale(r[ui])
            except ValueError:
                continue
            per[r[ki].split("(")[0]][r[ii]][r[mi]] = v
        stages = collections.defaultdict(lambda: {"dram_bytes": 0, "us": 0.0, "kernels": {}})
        for k, launches in per.items():
            name = k.replace("void ", "")
            st = next((s for p, s in STAGE if name.startswith(p)), None)
            if st is None:
                continue
            b = float(np.median([m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0) for m in launches.values()]))
            t = float(np.median([m.get("gpu__time_duration.sum", 0) for m in launches.values()]))
            stages[st]["kernels"][name] = {"dram_bytes": int(b), "us": round(t, 1), "launches": len(launches)}
            stages[st]["dram_bytes"] += int(b)
            stages[st]["us"] = round(stages[st]["us"] + t, 1)
        out[wl] = dict(stages, source=path)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
