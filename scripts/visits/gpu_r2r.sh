#!/bin/bash
# round 2, visit r (1 GPU): grid upload in one pipelined pass per pair of arrays (1 MiB slots), whole suite, e2e
TAG=${1:-r2r}
OUT=gpurun_out; mkdir -p $OUT
t0=$SECONDS
timeout 600 python -m pytest tests -m gpu -q -x > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$? ($((SECONDS-t0)) s)"; tail -2 $OUT/pytest_$TAG.log
GZB_CREATE_TIMING=2 timeout 120 python scripts/create_probe.py > $OUT/create_probe_$TAG.log 2>&1; echo "create probe rc=$? ($((SECONDS-t0)) s)"; grep "host threads\|cores" $OUT/create_probe_$TAG.log; grep "gzb create" $OUT/create_probe_$TAG.log | tail -5
timeout 120 python scripts/create_probe.py > $OUT/create_probe2_$TAG.log 2>&1; grep "host threads" $OUT/create_probe2_$TAG.log
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-f2 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$? ($((SECONDS-t0)) s)"; tail -2 $OUT/bench_$TAG.err
python - <<PY
import json
b=json.load(open("$OUT/bench_$TAG.json")); print(b["value"], b["ms_per_step"], b["e2e"]["ms_per_step"], b["e2e"]["breakdown_s"], b["details"]["k1_points_left_to_the_warp_search"], b["roofline"]["frac"], b["roofline"]["gather_kernel"]["frac"])
PY
echo "done ($((SECONDS-t0)) s)"
