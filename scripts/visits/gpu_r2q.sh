#!/bin/bash
# round 2, visit q (1 GPU): one shared-memory carve-out for the kernels of a step (option carveout), with and without persistent bulk blocks
TAG=${1:-r2q}
OUT=gpurun_out; mkdir -p $OUT
t0=$SECONDS
restore() { unset GZB_NVCC_EXTRA; python -m gonzag_b200.build --force > $OUT/build_restore_$TAG.log 2>&1; echo "product build restored rc=$?"; }
trap restore EXIT
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py tests/test_gpu_fuzz.py -m gpu -q -x > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$? ($((SECONDS-t0)) s)"; tail -2 $OUT/pytest_$TAG.log
AB_STAGES=K1,bulk,path timeout 300 python scripts/ab_opt.py carveout=-1,bulk_resident=0 carveout=28 carveout=14 carveout=44 carveout=28,bulk_resident=5 bulk_resident=6 bulk_resident=7 carveout=-1,bulk_resident=0 carveout=28 carveout=28,bulk_resident=6 carveout=28,bulk_resident=0,chain_first=0 > $OUT/ab_$TAG.log 2>&1; echo "ab rc=$? ($((SECONDS-t0)) s)"; cat $OUT/ab_$TAG.log
GZB_NVCC_EXTRA="-DGZB_KTIME=1" python -m gonzag_b200.build --force > $OUT/build_ktime_$TAG.log 2>&1; echo "ktime build rc=$?"
GZB_NVCC_EXTRA="-DGZB_KTIME=1" timeout 200 python scripts/ktime_case.py 3 carveout=28 > $OUT/ktime_cv28_$TAG.log 2>&1; echo "ktime rc=$? ($((SECONDS-t0)) s)"
GZB_NVCC_EXTRA="-DGZB_KTIME=1" timeout 200 python scripts/ktime_case.py 3 carveout=28 bulk_resident=6 > $OUT/ktime_cv28_res6_$TAG.log 2>&1; echo "ktime (bulk_resident=6) rc=$? ($((SECONDS-t0)) s)"
grep -h "step\|k_near_slow\|bulk kernel\|k_path_fix\|k_tail (- | end" $OUT/ktime_cv28_$TAG.log $OUT/ktime_cv28_res6_$TAG.log
echo "done ($((SECONDS-t0)) s)"
