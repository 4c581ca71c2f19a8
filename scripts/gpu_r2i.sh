#!/bin/bash
# round 2, visit i (1 GPU): adaptive window shapes of the warp search: parity, timings, launch lists, e2e
TAG=${1:-r2i}
OUT=gpurun_out; mkdir -p $OUT
t0=$SECONDS
timeout 500 python -m pytest tests -m gpu -q -x --durations=4 > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$? ($((SECONDS-t0)) s)"; tail -3 $OUT/pytest_$TAG.log
for WL in C3 C3saral C5w; do
  timeout 120 python scripts/prof_path.py $WL 3 > $OUT/prof_${WL}_$TAG.log 2>&1; echo "== $WL rc=$? ($((SECONDS-t0)) s)"; tail -3 $OUT/prof_${WL}_$TAG.log
done
for WL in C3 C3saral; do
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $OUT/launches_${WL}_$TAG.csv python scripts/prof_path.py $WL 2 > $OUT/ncu_${WL}_$TAG.log 2>&1; echo "ncu rc=$? ($((SECONDS-t0)) s)"
python scripts/launches.py $OUT/launches_${WL}_$TAG.csv | tail -7
done
python scripts/launches.py $OUT/launches_C3_$TAG.csv | grep -i "cellcert\|k_cert\|leaves"
AB_STAGES=K1,path timeout 90 python scripts/ab_opt.py pdl=1 > $OUT/ab_$TAG.log 2>&1; cat $OUT/ab_$TAG.log
timeout 200 python bench.py --no-cpu > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$? ($((SECONDS-t0)) s)"; tail -2 $OUT/bench_$TAG.err
python - <<'PY'
import json
b=json.load(open("gpurun_out/bench_r2i.json")); print(b["ms_per_step"], b["e2e"]["ms_per_step"], b["e2e"]["breakdown_s"], b["e2e"].get("resident_grid",{}).get("ms_per_step"), b["e2e"].get("with_xnptrack",{}).get("ms_per_step"))
PY
echo "done ($((SECONDS-t0)) s)"
