#!/bin/bash
# Short round-end check on one B200: GPU tests, smoke and the default bench line (with its genome / file blocks);
# NCU=1 adds the ncu launch list of a 3-step run.
out=gpurun_out/${TAG:-r2m}; mkdir -p $out
t0=$SECONDS
python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; tail -3 $out/pytest_gpu.log; echo "pytest $((SECONDS-t0)) s"
python __graft_entry__.py smoke > $out/smoke.log 2>&1; tail -1 $out/smoke.log
t0=$SECONDS
python bench.py > $out/bench_c2_n1.json 2> $out/bench_c2_n1.err || { tail -5 $out/bench_c2_n1.err; exit 1; }
echo "bench $((SECONDS-t0)) s"
if [ -n "$NCU" ]; then
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --genome none --no-file-e2e > $out/plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --genome none --no-file-e2e > $out/ncu_launches.log 2>&1
fi
python - <<'PY'
import json, os
j = json.loads(open("gpurun_out/%s/bench_c2_n1.json" % os.environ.get("TAG", "r2m")).read().strip().splitlines()[-1])
print("value %.4g ms_per_step %.4f e2e %s" % (j["value"], j["ms_per_step"], (j.get("e2e") or {}).get("ms_per_step")))
print("stage", j.get("stage_ms"))
print("file", {k: (v.get("ms"), v.get("read_ms")) for k, v in j.get("e2e_file", {}).items() if isinstance(v, dict)})
print("genome c3", {k: v.get("ms_per_pass") for k, v in j["genome"]["c3"]["by_columns"].items()})
PY
