#!/bin/bash
# round 2, visit w (1 GPU): room for the chain kernels on every SM -- resident blocks of the bulk kernel x register cap of k_chain
TAG=${1:-r2w}
OUT=gpurun_out; mkdir -p $OUT
t0=$SECONDS
restore() { unset GZB_NVCC_EXTRA; python -m gonzag_b200.build --force > $OUT/build_restore_$TAG.log 2>&1; echo "product build restored rc=$?"; }
trap restore EXIT
AB_STAGES=K1,bulk,path timeout 100 python scripts/ab_opt.py chain_first=6 chain_first=6 > $OUT/ab_base_$TAG.log 2>&1; echo "== product"; cat $OUT/ab_base_$TAG.log
k=0
for EXP in "-DGZB_CHAIN_MINBLOCKS=4" "-DGZB_EXP_BULK_BLOCKS=7 -DGZB_CHAIN_MINBLOCKS=4" "-DGZB_EXP_BULK_BLOCKS=6 -DGZB_CHAIN_MINBLOCKS=4" "-DGZB_EXP_BULK_BLOCKS=6" "-DGZB_EXP_BULK_BLOCKS=5 -DGZB_CHAIN_MINBLOCKS=4"; do
    k=$((k+1))
    echo "== experiment $k: $EXP ($((SECONDS-t0)) s)"
    GZB_NVCC_EXTRA="$EXP" python -m gonzag_b200.build --force > $OUT/build_exp${k}_$TAG.log 2>&1 || { echo "build failed"; continue; }
    AB_STAGES=K1,bulk,path GZB_NVCC_EXTRA="$EXP" timeout 90 python scripts/ab_opt.py chain_first=6 chain_first=6 > $OUT/ab_exp${k}_$TAG.log 2>&1; cat $OUT/ab_exp${k}_$TAG.log
done
echo "done ($((SECONDS-t0)) s)"
