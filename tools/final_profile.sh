#!/bin/bash
# Round-end measurement pass on one B200: GPU tests, the bench lines of every workload, the ncu launch list
# and one full capture of the dominant kernel (each ncu pass only after the plain run exited 0).
out=gpurun_out/${TAG:-r2}; mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; tail -3 $out/pytest_gpu.log
python bench.py > $out/bench_c2_n1.json 2> $out/bench_c2_n1.err || exit 1
python bench.py --workload c1 --steps 200 --no-cpu-baseline --genome none --no-file-e2e > $out/bench_c1_n1.json 2> $out/bench_c1_n1.err
python bench.py --workload c4 --steps 30 --no-cpu-baseline --genome none --no-file-e2e > $out/bench_c4_n1.json 2> $out/bench_c4_n1.err
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --genome none --no-file-e2e > $out/plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --genome none --no-file-e2e > $out/ncu_launches.log 2>&1
for w in c2 c4; do
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/traffic_$w.csv \
    python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline --genome none --no-file-e2e > $out/ncu_traffic_$w.log 2>&1
done
python tools/traffic.py c2=$out/traffic_c2.csv c4=$out/traffic_c4.csv > $out/roofline_traffic.json
ncu --set full --clock-control none --import-source on -k regex:k_score_fused -s 4 -c 1 -f -o $out/fused \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --genome none --no-file-e2e > $out/ncu_full.log 2>&1
ncu -i $out/fused.ncu-rep --page raw --csv > $out/fused_raw.csv 2>/dev/null
ncu -i $out/fused.ncu-rep --page source --csv > $out/fused_src.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:k_score_fused -s 2 -c 1 -f -o $out/fused_c4 \
    python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu-baseline --genome none --no-file-e2e > $out/ncu_full_c4.log 2>&1
ncu -i $out/fused_c4.ncu-rep --page raw --csv > $out/fused_c4_raw.csv 2>/dev/null
ncu -i $out/fused_c4.ncu-rep --page source --csv > $out/fused_c4_src.csv 2>/dev/null
python - <<'PY'
import json, os
for w in ("c2", "c1", "c4"):
    try:
        j = json.loads(open("gpurun_out/%s/bench_%s_n1.json" % (os.environ.get("TAG", "r2"), w)).read().strip().splitlines()[-1])
        print(w, "value %.4g ms_per_step %.4f e2e %s stage %s" % (j["value"], j["ms_per_step"], (j.get("e2e") or {}).get("ms_per_step"), j.get("stage_ms")))
    except Exception as e:
        print(w, "ERR", e)
PY
ls -la $out
