#!/bin/bash
# end-of-round evidence at HEAD (1 GPU): whole GPU suite, smoke, default bench (both legs), reference arm, ncu launch list of the bench
# command, one `--set full` capture of a step's kernels + the stand-alone stage kernels (exports only: the report stays on the box)
TAG=${1:-r2z}
OUT=gpurun_out; mkdir -p $OUT
t0=$SECONDS
timeout 600 python -m pytest tests -m gpu -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$? ($((SECONDS-t0)) s)"; tail -2 $OUT/pytest_$TAG.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke_$TAG.log 2>&1; echo "smoke rc=$? ($((SECONDS-t0)) s)"; tail -1 $OUT/smoke_$TAG.log
timeout 400 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$? ($((SECONDS-t0)) s)"; tail -2 $OUT/bench_$TAG.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "ref rc=$? ($((SECONDS-t0)) s)"; cut -c1-400 $OUT/bench_ref_$TAG.json
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k 'regex:^k_' -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-f2 > $OUT/ncu_launch_$TAG.log 2>&1; echo "ncu list rc=$? ($((SECONDS-t0)) s)"
timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on -o $OUT/prof_$TAG -f \
    python scripts/ncu_case.py > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$? ($((SECONDS-t0)) s)"; tail -2 $OUT/ncu_full_$TAG.log
if [ -f $OUT/prof_$TAG.ncu-rep ]; then
  ncu -i $OUT/prof_$TAG.ncu-rep --page raw --csv > $OUT/prof_${TAG}_raw.csv 2>/dev/null
  ls -la $OUT/prof_$TAG.ncu-rep
  sz=$(stat -c %s $OUT/prof_$TAG.ncu-rep); if [ $sz -gt 40000000 ]; then rm -f $OUT/prof_$TAG.ncu-rep; echo "report dropped (too large to bring back), CSV export kept"; fi
fi
python - <<PY
import json
b=json.load(open("$OUT/bench_$TAG.json")); e=b["e2e"]
print("value %.4g pts/s  ms/step %.4f  e2e %.2f ms (%.4g pts/s)  roofline %.3f  gather %.3f  cpu %s" % (b["value"], b["ms_per_step"], e["ms_per_step"], e["value"], b["roofline"]["frac"], b["roofline"]["gather_kernel"]["frac"], b["cpu_baseline"]))
PY
echo "done ($((SECONDS-t0)) s)"
