#!/bin/bash
# round 2, visit b (1 GPU): new GPU tests, default bench, C5s windowed arm, reference arm smoke
TAG=${1:-r2b}
OUT=gpurun_out; mkdir -p $OUT
t0=$SECONDS
timeout 400 python -m pytest tests -m gpu -q -x --durations=8 --deselect tests/test_gpu_fullsize.py > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$? ($((SECONDS-t0)) s)"; tail -4 $OUT/pytest_$TAG.log
timeout 200 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$? ($((SECONDS-t0)) s)"; tail -3 $OUT/bench_$TAG.err
timeout 200 python bench.py --workload C5s --steps 2 --warmup 1 > $OUT/bench_c5s_$TAG.json 2> $OUT/bench_c5s_$TAG.err; echo "bench C5s rc=$? ($((SECONDS-t0)) s)"; tail -3 $OUT/bench_c5s_$TAG.err; cut -c1-600 $OUT/bench_c5s_$TAG.json
timeout 200 python bench.py --impl reference --steps 1 --warmup 0 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "bench ref rc=$? ($((SECONDS-t0)) s)"; tail -2 $OUT/bench_ref_$TAG.err; cut -c1-900 $OUT/bench_ref_$TAG.json
timeout 120 python -m pytest tests/test_oracle_vs_reference.py -q -x > $OUT/pytest_ref_$TAG.log 2>&1; echo "ref tests rc=$? ($((SECONDS-t0)) s)"; tail -2 $OUT/pytest_ref_$TAG.log
echo "done ($((SECONDS-t0)) s)"
