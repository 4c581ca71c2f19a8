#!/bin/bash
# round 2, visit j (1 GPU): gather variants (64-byte L2 fetches, early neighbourhood prefetch, register cap), sanitizers
TAG=${1:-r2j}
OUT=gpurun_out; mkdir -p $OUT
t0=$SECONDS
restore() { unset GZB_NVCC_EXTRA; python -m gonzag_b200.build --force > $OUT/build_restore_$TAG.log 2>&1; echo "product build restored rc=$?"; }
timeout 300 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$? ($((SECONDS-t0)) s)"; tail -2 $OUT/pytest_$TAG.log
AB_STAGES=K4,bulk,path timeout 120 python scripts/ab_opt.py prefetch=0 prefetch=3 prefetch=4 prefetch=0 prefetch=4 > $OUT/ab_$TAG.log 2>&1; cat $OUT/ab_$TAG.log
for TOOL in memcheck racecheck synccheck; do
  timeout 400 compute-sanitizer --tool $TOOL python scripts/san_case.py > $OUT/sanitizer_${TOOL}_$TAG.log 2>&1; echo "sanitizer $TOOL rc=$? ($((SECONDS-t0)) s)"; tail -4 $OUT/sanitizer_${TOOL}_$TAG.log
done
trap restore EXIT
for EXP in "-DGZB_EXP_BULK_BLOCKS=7" "-DGZB_EXP_BULK_BLOCKS=6"; do
    echo "== experiment: $EXP ($((SECONDS-t0)) s)"
    GZB_NVCC_EXTRA="$EXP" python -m gonzag_b200.build --force > $OUT/build_exp_$TAG.log 2>&1 || { echo "build failed"; tail -5 $OUT/build_exp_$TAG.log; continue; }
    AB_STAGES=bulk,path GZB_NVCC_EXTRA="$EXP" timeout 90 python scripts/ab_opt.py prefetch=0 prefetch=4 prefetch=3
done
echo "done ($((SECONDS-t0)) s)"
