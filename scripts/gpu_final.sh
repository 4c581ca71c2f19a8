#!/bin/bash
# What the driver runs at round end, in one visit: GPU tests, smoke(), both bench arms; plus the K4 stand-alone ncu capture.
TAG=${1:-final}
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 $OUT/smoke_$TAG.log
python bench.py --impl reference --steps 1 --warmup 0 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "ref rc=$?"
python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"; tail -2 $OUT/bench_$TAG.err
ncu --set full --clock-control none --import-source on -k 'regex:^k_interp' --launch-skip 1 --launch-count 1 -o $OUT/prof_${TAG}_k4 -f \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_full_${TAG}_k4.log 2>&1; echo "ncu full (K4) rc=$?"
python - <<PY
import json
d=json.load(open("$OUT/bench_$TAG.json"))
print("value %.4g pts/s  ms/step %.4f  launches %d  clocks %s" % (d["value"], d["ms_per_step"], d["gpu_launches"], d["clocks"]))
print("roofline", d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"]["traffic"])
print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["breakdown_s"])
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
r=json.load(open("$OUT/bench_ref_$TAG.json")); print("reference arm", r["value"], r["cpu_baseline"]["cores"])
PY
