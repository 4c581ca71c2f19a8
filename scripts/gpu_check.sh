#!/bin/bash
# One short visit to a B200 box, most important first, every part under its own time limit:
#   GPU tests (without the full-size ones), bench.py, smoke(), ncu launch list, full-size tests.
#   bash scripts/gpu_check.sh TAG
TAG=${1:-chk}
OUT=gpurun_out; mkdir -p $OUT
t0=$SECONDS
timeout 260 python -m pytest tests -m gpu -q --durations=12 --deselect tests/test_gpu_fullsize.py > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$? ($((SECONDS-t0)) s)"; tail -4 $OUT/pytest_$TAG.log
timeout 170 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$? ($((SECONDS-t0)) s)"; tail -2 $OUT/bench_$TAG.err
python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$TAG.json"))
    print("value %.4g pts/s  ms/step %.4f  launches %d  clocks %s" % (d["value"], d["ms_per_step"], d["gpu_launches"], d["clocks"]))
    print("roofline", json.dumps(d["roofline"])[:600])
    for k,v in d["kernels"].items():
        if isinstance(v, dict): print("  %-18s %.4f ms (min %.4f)  %.0f GB/s  frac %.3f" % (k, v["ms"], v.get("ms_min",0), v["achieved_gbs"], v["frac_of_hbm_peak"]))
    print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"]); print("cpu", d["cpu_baseline"]["value"])
except Exception as e: print("bench line unreadable:", e)
PY
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$? ($((SECONDS-t0)) s)"; tail -1 $OUT/smoke_$TAG.log
timeout 150 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k 'regex:^k_' -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_launch_$TAG.log 2>&1; echo "ncu list rc=$? ($((SECONDS-t0)) s)"
timeout 400 python -m pytest tests/test_gpu_fullsize.py -m gpu -q --durations=8 > $OUT/pytest_full_$TAG.log 2>&1; echo "pytest fullsize rc=$? ($((SECONDS-t0)) s)"; tail -3 $OUT/pytest_full_$TAG.log
