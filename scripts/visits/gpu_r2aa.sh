#!/bin/bash
# round 2, visit aa (1 GPU): exact float32 narrowing of the lat / lon upload
OUT=gpurun_out; mkdir -p $OUT
t0=$SECONDS
timeout 400 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_dropin.py tests/test_gpu_edge.py -m gpu -q -x > $OUT/pytest_r2aa.log 2>&1; echo "pytest rc=$? ($((SECONDS-t0)) s)"; tail -3 $OUT/pytest_r2aa.log
GZB_CREATE_TIMING=1 timeout 120 python scripts/create_probe.py > $OUT/create_probe_r2aa.log 2>&1; grep "host threads" $OUT/create_probe_r2aa.log | head -3; grep "gzb create" $OUT/create_probe_r2aa.log | tail -4
timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu --no-f2 > $OUT/bench_r2aa.json 2> $OUT/bench_r2aa.err; echo "bench rc=$? ($((SECONDS-t0)) s)"; tail -1 $OUT/bench_r2aa.err
python - <<PY
import json
b=json.load(open("$OUT/bench_r2aa.json")); e=b["e2e"]; print(b["ms_per_step"], e["ms_per_step"], e["h2d_bytes_per_step"], e["breakdown_s"], e["resident_grid"]["ms_per_step"])
PY
echo "done ($((SECONDS-t0)) s)"
