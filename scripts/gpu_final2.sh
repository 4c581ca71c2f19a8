#!/bin/bash
# end-of-round numbers: default bench (both legs), smoke, reference arm
TAG=${1:-fin}
OUT=gpurun_out; mkdir -p $OUT
t0=$SECONDS
timeout 200 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$? ($((SECONDS-t0)) s)"; tail -2 $OUT/bench_$TAG.err
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$TAG.json"))
    print("value %.4g pts/s  ms/step %.4f  launches %d  clocks %s" % (d["value"], d["ms_per_step"], d["gpu_launches"], d["clocks"]))
    print("roofline", json.dumps({k:v for k,v in d["roofline"].items() if k!="note"}))
    for k,v in d["kernels"].items():
        if isinstance(v, dict): print("  %-18s %.4f ms (min %.4f)  %.0f GB/s  frac %.3f" % (k, v["ms"], v.get("ms_min",0), v["achieved_gbs"], v["frac_of_hbm_peak"]))
    print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["breakdown_s"]); print("cpu", d["cpu_baseline"])
except Exception as e: print("bench line unreadable:", e)
PY
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$? ($((SECONDS-t0)) s)"; tail -1 $OUT/smoke_$TAG.log
timeout 120 python bench.py --impl reference --steps 1 --warmup 0 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "ref rc=$? ($((SECONDS-t0)) s)"; cut -c1-300 $OUT/bench_ref_$TAG.json
