#!/bin/bash
# round 2, visit e (1 GPU): per-kernel launch lists (ncu, serialised) of a whole step: C3, polar C3, 3-mission window
TAG=${1:-r2e}
OUT=gpurun_out; mkdir -p $OUT
t0=$SECONDS
for WL in C3 C3saral C5w; do
  timeout 120 python scripts/prof_path.py $WL 3 > $OUT/prof_${WL}_$TAG.log 2>&1; echo "== $WL rc=$? ($((SECONDS-t0)) s)"; tail -4 $OUT/prof_${WL}_$TAG.log
  timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $OUT/launches_${WL}_$TAG.csv python scripts/prof_path.py $WL 2 > $OUT/ncu_${WL}_$TAG.log 2>&1; echo "ncu rc=$? ($((SECONDS-t0)) s)"
  python scripts/launches.py $OUT/launches_${WL}_$TAG.csv | tail -12
done
echo "done ($((SECONDS-t0)) s)"
