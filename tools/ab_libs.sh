#!/bin/bash
# A/B of library builds / tuning flags on one box:
#   tools/ab_libs.sh <workload> <steps> <name=path[,bench flags]>...
# PEAKACHU_B200_LIB selects the build; prints ms per step, fused-kernel ms and e2e ms per variant, two repeats.
wl=$1; steps=$2; shift 2
mkdir -p gpurun_out/ab
for rep in 1 2; do
for kv in "$@"; do
  name=${kv%%=*}; rest=${kv#*=}; path=${rest%%,*}; flags=""
  if [ "$rest" != "$path" ]; then flags=${rest#*,}; fi
  PEAKACHU_B200_LIB=$path python bench.py --workload $wl --steps $steps --warmup 3 --no-cpu-baseline $flags > gpurun_out/ab/${wl}_${name}_$rep.json 2> gpurun_out/ab/${wl}_${name}_$rep.err
  python - "$wl" "$name" "$rep" <<'PY'
import json, sys
wl, name, rep = sys.argv[1:4]
try:
    j = json.loads(open("gpurun_out/ab/%s_%s_%s.json" % (wl, name, rep)).read().strip().splitlines()[-1])
    print(wl, name, rep, "ms_per_step %.4f fused_ms %.4f e2e_ms %.4f" % (j["ms_per_step"], j["stage_ms"]["features"], j["e2e"]["ms_per_step"]))
except Exception as e:
    print(wl, name, rep, "ERR", e)
PY
done
done
