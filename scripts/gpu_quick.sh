#!/bin/bash
# quick GPU check: parity tests + bench without the CPU legs.  bash scripts/gpu_quick.sh TAG [bench args]
TAG=${1:-q}; shift
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 $OUT/pytest_$TAG.log
python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e "$@" > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; tail -3 $OUT/bench_$TAG.err
python - <<PY
import json
d=json.load(open("$OUT/bench_$TAG.json"))
print("value %.4g pts/s  ms/step %.4f  launches %d" % (d["value"], d["ms_per_step"], d["gpu_launches"]))
for k,v in d["kernels"].items():
    if isinstance(v, dict): print("  %-18s %.4f ms (min %.4f)  %.0f GB/s  frac %.3f" % (k, v["ms"], v.get("ms_min",0), v["achieved_gbs"], v["frac_of_hbm_peak"]))
PY
