#!/bin/bash
# round 2, visit p (1 GPU): K1 tail (k_tail write-back from shared memory, k_chain ordered copy in one pass), release order of the
# bulk kernel (chain_first), persistent bulk blocks (bulk_resident), copy threads of the grid upload
TAG=${1:-r2p}
OUT=gpurun_out; mkdir -p $OUT
t0=$SECONDS
restore() { unset GZB_NVCC_EXTRA; python -m gonzag_b200.build --force > $OUT/build_restore_$TAG.log 2>&1; echo "product build restored rc=$?"; }
trap restore EXIT
timeout 600 python -m pytest tests -m gpu -q -x > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$? ($((SECONDS-t0)) s)"; tail -3 $OUT/pytest_$TAG.log
AB_STAGES=K1,bulk,path timeout 300 python scripts/ab_opt.py chain_first=0,bulk_resident=0 chain_first=1 bulk_resident=4 bulk_resident=5 bulk_resident=6 bulk_resident=7 bulk_resident=0 chain_first=0 chain_first=1,bulk_resident=5 bulk_resident=6 bulk_resident=0 > $OUT/ab_$TAG.log 2>&1; echo "ab rc=$? ($((SECONDS-t0)) s)"; cat $OUT/ab_$TAG.log
timeout 120 python scripts/create_probe.py > $OUT/create_probe_$TAG.log 2>&1; echo "create probe rc=$? ($((SECONDS-t0)) s)"; cat $OUT/create_probe_$TAG.log
GZB_NVCC_EXTRA="-DGZB_KTIME=1" python -m gonzag_b200.build --force > $OUT/build_ktime_$TAG.log 2>&1; echo "ktime build rc=$?"
GZB_NVCC_EXTRA="-DGZB_KTIME=1" timeout 200 python scripts/ktime_case.py 3 > $OUT/ktime_$TAG.log 2>&1; echo "ktime rc=$? ($((SECONDS-t0)) s)"
GZB_NVCC_EXTRA="-DGZB_KTIME=1" timeout 200 python scripts/ktime_case.py 3 bulk_resident=5 > $OUT/ktime_res5_$TAG.log 2>&1; echo "ktime (bulk_resident=5) rc=$? ($((SECONDS-t0)) s)"
echo "done ($((SECONDS-t0)) s)"
