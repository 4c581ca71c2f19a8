#!/bin/bash
# round 2, visit c (1 GPU): fused K1 tail + multi-chain call: GPU tests, stage timings, default bench, C5s windowed arm, stream trace
TAG=${1:-r2c}
OUT=gpurun_out; mkdir -p $OUT
t0=$SECONDS
timeout 500 python -m pytest tests -m gpu -q -x --durations=6 --deselect tests/test_gpu_fullsize.py > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$? ($((SECONDS-t0)) s)"; tail -4 $OUT/pytest_$TAG.log
AB_STAGES=K4,bulk,K2K3,K1,path timeout 90 python scripts/ab_opt.py pdl=1 > $OUT/ab_$TAG.log 2>&1; cat $OUT/ab_$TAG.log
timeout 200 python bench.py --no-cpu > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$? ($((SECONDS-t0)) s)"; tail -3 $OUT/bench_$TAG.err
GZB_TRACE_STEP=1 timeout 120 python bench.py --no-cpu --no-e2e --steps 2 > /dev/null 2> $OUT/trace_$TAG.err; grep -A40 -i "trace\|K1 start" $OUT/trace_$TAG.err | head -60
timeout 200 python bench.py --workload C5s --steps 2 --warmup 1 > $OUT/bench_c5s_$TAG.json 2> $OUT/bench_c5s_$TAG.err; echo "bench C5s rc=$? ($((SECONDS-t0)) s)"; tail -3 $OUT/bench_c5s_$TAG.err; cut -c1-300 $OUT/bench_c5s_$TAG.json
timeout 300 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -x > $OUT/pytest_full_$TAG.log 2>&1; echo "fullsize rc=$? ($((SECONDS-t0)) s)"; tail -3 $OUT/pytest_full_$TAG.log
echo "done ($((SECONDS-t0)) s)"
