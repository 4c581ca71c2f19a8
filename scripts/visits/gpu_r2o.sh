#!/bin/bash
# round 2, visit o (1 GPU): leaner k_cellcert (same certificates expected: the count of points left to the warp search must
# stay 4173 on C3), whole GPU suite, where a context's milliseconds go (GZB_CREATE_TIMING), in-kernel time marks of a step (KTIME build)
TAG=${1:-r2o}
OUT=gpurun_out; mkdir -p $OUT
t0=$SECONDS
restore() { unset GZB_NVCC_EXTRA; python -m gonzag_b200.build --force > $OUT/build_restore_$TAG.log 2>&1; echo "product build restored rc=$?"; }
trap restore EXIT
timeout 600 python -m pytest tests -m gpu -q -x > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$? ($((SECONDS-t0)) s)"; tail -3 $OUT/pytest_$TAG.log
GZB_CREATE_TIMING=1 timeout 200 python scripts/e2e_phases.py 3 > $OUT/e2e_phases_$TAG.log 2>&1; echo "e2e phases rc=$? ($((SECONDS-t0)) s)"; grep -A6 "rep 2" $OUT/e2e_phases_$TAG.log | tail -12; grep "gzb create" $OUT/e2e_phases_$TAG.log | tail -12
GZB_CREATE_TIMING=2 timeout 200 python scripts/e2e_phases.py 2 > $OUT/e2e_phases2_$TAG.log 2>&1; grep "gzb create" $OUT/e2e_phases2_$TAG.log | tail -6
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-f2 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$? ($((SECONDS-t0)) s)"; tail -2 $OUT/bench_$TAG.err
python - <<PY
import json
b=json.load(open("$OUT/bench_$TAG.json")); print(b["value"], b["ms_per_step"], b["e2e"]["ms_per_step"], b["e2e"]["breakdown_s"], b["details"]["k1_points_left_to_the_warp_search"], b["roofline"]["frac"], b["roofline"]["gather_kernel"]["frac"])
PY
GZB_NVCC_EXTRA="-DGZB_KTIME=1" python -m gonzag_b200.build --force > $OUT/build_ktime_$TAG.log 2>&1; echo "ktime build rc=$?"
GZB_NVCC_EXTRA="-DGZB_KTIME=1" timeout 200 python scripts/ktime_case.py 4 > $OUT/ktime_$TAG.log 2>&1; echo "ktime rc=$? ($((SECONDS-t0)) s)"; tail -60 $OUT/ktime_$TAG.log
echo "done ($((SECONDS-t0)) s)"
