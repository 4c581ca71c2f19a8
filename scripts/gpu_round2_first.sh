#!/bin/bash
# First GPU visit of a new round (~15 min of box time: ~5 for the green check + fuzz, ~1.5 per experiment): is everything still green, and what do the compile-time
# experiment switches of csrc/gzb_common.cuh buy?  Each experiment: forced rebuild with GZB_NVCC_EXTRA (13 s),
# stage timings on the C3 workload (scripts/ab_opt.py: K4, bulk K2K3K4, K2K3, whole path), parity tests on the
# golden cases.  The product build is restored at the end whatever happens.
#   bash scripts/gpu_round2_first.sh TAG
TAG=${1:-r2a}
OUT=gpurun_out; mkdir -p $OUT
t0=$SECONDS
restore() { unset GZB_NVCC_EXTRA; python -m gonzag_b200.build --force > $OUT/build_restore_$TAG.log 2>&1; echo "product build restored rc=$?"; }
trap restore EXIT
timeout 260 python -m pytest tests -m gpu -q --durations=8 --deselect tests/test_gpu_fullsize.py > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$? ($((SECONDS-t0)) s)"; tail -3 $OUT/pytest_$TAG.log
timeout 170 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$? ($((SECONDS-t0)) s)"; tail -2 $OUT/bench_$TAG.err
timeout 120 python scripts/gpu_fuzz.py 5000 600 > $OUT/fuzz_$TAG.log 2>&1; echo "fuzz rc=$? ($((SECONDS-t0)) s)"; tail -5 $OUT/fuzz_$TAG.log
AB_STAGES=K4,bulk,K2K3,K1,path timeout 60 python scripts/ab_opt.py pdl=1 > $OUT/ab_base_$TAG.log 2>&1; echo "== product build"; cat $OUT/ab_base_$TAG.log
k=0
for EXP in "-DGZB_EXP_SKIP_DG=1" "-DGZB_EXP_L64=1" "-DGZB_EXP_HDG_STRIDE=8" "-DGZB_EXP_L64=1 -DGZB_EXP_HDG_STRIDE=8" "-DGZB_EXP_BULK_BLOCKS=6" "-DGZB_EXP_BULK_BLOCKS=7" \
           "-DGZB_EXP_L64=1 -DGZB_EXP_HDG_STRIDE=8 -DGZB_EXP_BULK_BLOCKS=6"; do
    k=$((k+1))
    echo "== experiment $k: $EXP ($((SECONDS-t0)) s)"
    GZB_NVCC_EXTRA="$EXP" python -m gonzag_b200.build --force > $OUT/build_exp${k}_$TAG.log 2>&1 || { echo "build failed"; tail -5 $OUT/build_exp${k}_$TAG.log; continue; }
    grep -h "k_mesh_weights_interpIfLi0\|Used" $OUT/build_exp${k}_$TAG.log | grep -A1 "k_mesh_weights_interpIfLi0" | tail -1
    AB_STAGES=K4,bulk,K2K3,K1,path GZB_NVCC_EXTRA="$EXP" timeout 60 python scripts/ab_opt.py pdl=1 > $OUT/ab_exp${k}_$TAG.log 2>&1; cat $OUT/ab_exp${k}_$TAG.log
    GZB_NVCC_EXTRA="$EXP" timeout 150 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py -m gpu -q -x -k "global or regional or seam or config_vs_oracle or overlap_convention or smallest or nonfinite" > $OUT/pytest_exp${k}_$TAG.log 2>&1; echo "parity rc=$?"; tail -1 $OUT/pytest_exp${k}_$TAG.log
done
echo "done ($((SECONDS-t0)) s)"
