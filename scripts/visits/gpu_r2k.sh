#!/bin/bash
# round 2, visit k (1 GPU): CHECKED build (own bounds checks) through the GPU tests; C5 and C4 at N = 1; default bench
TAG=${1:-r2k}
OUT=gpurun_out; mkdir -p $OUT
t0=$SECONDS
restore() { unset GZB_NVCC_EXTRA; python -m gonzag_b200.build --force > $OUT/build_restore_$TAG.log 2>&1; echo "product build restored rc=$?"; }
trap restore EXIT
GZB_NVCC_EXTRA="-DGZB_CHECKED=1" python -m gonzag_b200.build --force > $OUT/build_checked_$TAG.log 2>&1; echo "checked build rc=$?"
GZB_NVCC_EXTRA="-DGZB_CHECKED=1" timeout 600 python -m pytest tests -m gpu -q -x > $OUT/pytest_checked_$TAG.log 2>&1; echo "pytest (CHECKED build) rc=$? ($((SECONDS-t0)) s)"; tail -3 $OUT/pytest_checked_$TAG.log
GZB_NVCC_EXTRA="-DGZB_CHECKED=1" timeout 120 python scripts/san_case.py > $OUT/san_case_checked_$TAG.log 2>&1; echo "san_case (CHECKED build) rc=$?"; tail -1 $OUT/san_case_checked_$TAG.log
GZB_NVCC_EXTRA="-DGZB_CHECKED=1" python -c "import gonzag_b200._lib as L; print(L.lib().gzb_version().decode())"
restore; trap - EXIT
timeout 300 python bench.py --workload C5 --steps 3 --warmup 1 > $OUT/bench_c5_n1_$TAG.json 2> $OUT/bench_c5_n1_$TAG.err; echo "bench C5 N=1 rc=$? ($((SECONDS-t0)) s)"; tail -2 $OUT/bench_c5_n1_$TAG.err; cut -c1-260 $OUT/bench_c5_n1_$TAG.json
timeout 400 python bench.py --workload C4 --steps 3 --warmup 1 > $OUT/bench_c4_n1_$TAG.json 2> $OUT/bench_c4_n1_$TAG.err; echo "bench C4 N=1 rc=$? ($((SECONDS-t0)) s)"; tail -2 $OUT/bench_c4_n1_$TAG.err; cut -c1-260 $OUT/bench_c4_n1_$TAG.json
timeout 300 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$? ($((SECONDS-t0)) s)"; tail -2 $OUT/bench_$TAG.err
python - <<'PY'
import json
b=json.load(open("gpurun_out/bench_r2k.json")); print(b["value"], b["ms_per_step"], b["e2e"]["ms_per_step"], b["e2e"].get("streamed"), b["cpu_baseline"])
PY
echo "done ($((SECONDS-t0)) s)"
