#!/bin/bash
# round 2, visit n (1 GPU): 32-bit index variants of K2-K4 -- parity tests, A/B against the 64-bit variants (option ix64) on the C3
# workload, resident blocks per SM of the fused kernel and of K4 (compile-time), then the end-to-end phase probe.
TAG=${1:-r2n}
OUT=gpurun_out; mkdir -p $OUT
t0=$SECONDS
restore() { unset GZB_NVCC_EXTRA; python -m gonzag_b200.build --force > $OUT/build_restore_$TAG.log 2>&1; echo "product build restored rc=$?"; }
trap restore EXIT
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py tests/test_gpu_fuzz.py -m gpu -q -x > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$? ($((SECONDS-t0)) s)"; tail -3 $OUT/pytest_$TAG.log
AB_STAGES=K4,bulk,K2K3,path timeout 120 python scripts/ab_opt.py ix64=1 ix64=0 ix64=1 ix64=0 > $OUT/ab_ix_$TAG.log 2>&1; echo "== product build (BULK 7, K4 6): ix64 A/B"; cat $OUT/ab_ix_$TAG.log
k=0
for EXP in "-DGZB_EXP_BULK_BLOCKS=6 -DGZB_K4_BLOCKS=5" "-DGZB_EXP_BULK_BLOCKS=5 -DGZB_K4_BLOCKS=7" "-DGZB_EXP_BULK_BLOCKS=8 -DGZB_K4_BLOCKS=8"; do
    k=$((k+1))
    echo "== experiment $k: $EXP ($((SECONDS-t0)) s)"
    GZB_NVCC_EXTRA="$EXP" python -m gonzag_b200.build --force > $OUT/build_exp${k}_$TAG.log 2>&1 || { echo "build failed"; tail -5 $OUT/build_exp${k}_$TAG.log; continue; }
    AB_STAGES=K4,bulk,path GZB_NVCC_EXTRA="$EXP" timeout 90 python scripts/ab_opt.py ix64=0 ix64=0 > $OUT/ab_exp${k}_$TAG.log 2>&1; cat $OUT/ab_exp${k}_$TAG.log
done
restore; trap - EXIT
timeout 200 python scripts/e2e_phases.py 4 > $OUT/e2e_phases_$TAG.log 2>&1; echo "e2e phases rc=$? ($((SECONDS-t0)) s)"; tail -8 $OUT/e2e_phases_$TAG.log
echo "done ($((SECONDS-t0)) s)"
