#!/bin/bash
# round 2, visit f (1 GPU): weak-cell table (wide boxes + portals) in the per-thread search: parity, timings, launch lists
TAG=${1:-r2f}
OUT=gpurun_out; mkdir -p $OUT
t0=$SECONDS
restore() { unset GZB_NVCC_EXTRA; python -m gonzag_b200.build --force > $OUT/build_restore_$TAG.log 2>&1; echo "product build restored rc=$?"; }
timeout 500 python -m pytest tests -m gpu -q -x --durations=6 > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$? ($((SECONDS-t0)) s)"; tail -4 $OUT/pytest_$TAG.log
for WL in C3 C3saral C5w; do
  timeout 120 python scripts/prof_path.py $WL 3 > $OUT/prof_${WL}_$TAG.log 2>&1; echo "== $WL rc=$? ($((SECONDS-t0)) s)"; tail -3 $OUT/prof_${WL}_$TAG.log
done
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file $OUT/launches_C5w_$TAG.csv python scripts/prof_path.py C5w 2 > $OUT/ncu_C5w_$TAG.log 2>&1; echo "ncu rc=$? ($((SECONDS-t0)) s)"
python scripts/launches.py $OUT/launches_C5w_$TAG.csv | tail -7
python scripts/launches.py $OUT/launches_C5w_$TAG.csv | grep -i "weak\|cellcert\|k_cert\|lut\|leaves"
AB_STAGES=K1,path timeout 90 python scripts/ab_opt.py pdl=1 > $OUT/ab_$TAG.log 2>&1; cat $OUT/ab_$TAG.log
timeout 200 python bench.py --no-cpu > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$? ($((SECONDS-t0)) s)"; tail -2 $OUT/bench_$TAG.err
trap restore EXIT
for EXP in "-DGZB_CH_PPT=2" "-DGZB_CH_PPT=1"; do
    echo "== experiment: $EXP ($((SECONDS-t0)) s)"
    GZB_NVCC_EXTRA="$EXP" python -m gonzag_b200.build --force > $OUT/build_exp_$TAG.log 2>&1 || { echo "build failed"; tail -5 $OUT/build_exp_$TAG.log; continue; }
    AB_STAGES=K1,path GZB_NVCC_EXTRA="$EXP" timeout 60 python scripts/ab_opt.py pdl=1
done
echo "done ($((SECONDS-t0)) s)"
