#!/bin/bash
# round 2, visit m (1 GPU): green check at HEAD, default bench, ncu launch list of the bench command,
# one `--set full` capture (with source) of a step's kernels + the stand-alone stage kernels, exported to CSV on the box.
TAG=${1:-r2m}
OUT=gpurun_out; mkdir -p $OUT
t0=$SECONDS
timeout 600 python -m pytest tests -m gpu -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$? ($((SECONDS-t0)) s)"; tail -2 $OUT/pytest_$TAG.log
timeout 300 python bench.py > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$? ($((SECONDS-t0)) s)"; tail -2 $OUT/bench_$TAG.err
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k 'regex:^k_' -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-f2 > $OUT/ncu_launch_$TAG.log 2>&1; echo "ncu list rc=$? ($((SECONDS-t0)) s)"
timeout 500 ncu --profile-from-start off --set full --clock-control none --import-source on -o $OUT/prof_$TAG -f \
    python scripts/ncu_case.py > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$? ($((SECONDS-t0)) s)"; tail -2 $OUT/ncu_full_$TAG.log
if [ -f $OUT/prof_$TAG.ncu-rep ]; then
  ncu -i $OUT/prof_$TAG.ncu-rep --page raw --csv > $OUT/prof_${TAG}_raw.csv 2>/dev/null
  for K in k_mesh_weights_interp k_interp k_near_search k_near_slow k_chain k_tail k_path_fix; do
    ncu -i $OUT/prof_$TAG.ncu-rep --page source --csv --print-source cuda -k "regex:$K" > $OUT/src_${TAG}_$K.csv 2>/dev/null
    ncu -i $OUT/prof_$TAG.ncu-rep --page source --csv --print-source sass -k "regex:$K" 2>/dev/null | gzip > $OUT/sass_${TAG}_$K.csv.gz
  done
  ls -la $OUT/prof_$TAG.ncu-rep
  sz=$(stat -c %s $OUT/prof_$TAG.ncu-rep); if [ $sz -gt 30000000 ]; then rm -f $OUT/prof_$TAG.ncu-rep; echo "report dropped (too large to bring back), CSV exports kept"; fi
fi
du -sh $OUT | tail -1
echo "done ($((SECONDS-t0)) s)"
