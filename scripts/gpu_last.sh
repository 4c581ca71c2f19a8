#!/bin/bash
# last check of the round at HEAD (1 GPU): whole GPU suite, smoke, default bench
OUT=gpurun_out; mkdir -p $OUT
t0=$SECONDS
timeout 600 python -m pytest tests -m gpu -q > $OUT/pytest_last.log 2>&1; echo "pytest rc=$? ($((SECONDS-t0)) s)"; tail -2 $OUT/pytest_last.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_last.log 2>&1; echo "smoke rc=$? ($((SECONDS-t0)) s)"; tail -1 $OUT/smoke_last.log
timeout 300 python bench.py > $OUT/bench_last.json 2> $OUT/bench_last.err; echo "bench rc=$? ($((SECONDS-t0)) s)"; tail -1 $OUT/bench_last.err
python - <<PY
import json
b=json.load(open("$OUT/bench_last.json")); e=b["e2e"]
print("value %.4g pts/s  ms/step %.4f  e2e %.2f ms (%.4g pts/s)  roofline %.3f  gather %.3f  cpu %.1f pts/s (%s)  launches %d  clocks %s" % (b["value"], b["ms_per_step"], e["ms_per_step"], e["value"], b["roofline"]["frac"], b["roofline"]["gather_kernel"]["frac"], b["cpu_baseline"]["value"], b["cpu_baseline"]["kind"], b["gpu_launches"], b["clocks"]))
PY
echo "done ($((SECONDS-t0)) s)"
