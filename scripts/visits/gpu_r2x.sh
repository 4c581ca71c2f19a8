#!/bin/bash
# round 2, visit x (1 GPU): resident blocks of the warp search (k_near_slow) -- one wave for the 2368 blocks?
TAG=${1:-r2x}
OUT=gpurun_out; mkdir -p $OUT
t0=$SECONDS
restore() { unset GZB_NVCC_EXTRA; python -m gonzag_b200.build --force > $OUT/build_restore_$TAG.log 2>&1; echo "product build restored rc=$?"; }
trap restore EXIT
AB_STAGES=K1,path timeout 100 python scripts/ab_opt.py pdl=1 pdl=1 > $OUT/ab_base_$TAG.log 2>&1; echo "== product"; cat $OUT/ab_base_$TAG.log
k=0
for EXP in "-DGZB_SLOW_MINBLOCKS=12" "-DGZB_SLOW_MINBLOCKS=16" "-DGZB_SLOW_MINBLOCKS=8"; do
    k=$((k+1))
    echo "== experiment $k: $EXP ($((SECONDS-t0)) s)"
    GZB_NVCC_EXTRA="$EXP" python -m gonzag_b200.build --force > $OUT/build_exp${k}_$TAG.log 2>&1 || { echo "build failed"; continue; }
    AB_STAGES=K1,path GZB_NVCC_EXTRA="$EXP" timeout 90 python scripts/ab_opt.py pdl=1 pdl=1 > $OUT/ab_exp${k}_$TAG.log 2>&1; cat $OUT/ab_exp${k}_$TAG.log
done
echo "done ($((SECONDS-t0)) s)"
