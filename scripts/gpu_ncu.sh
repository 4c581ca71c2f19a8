#!/bin/bash
# ncu launch list (+ optional --set full capture of the step's kernels) of the bench workload.
#   bash scripts/gpu_ncu.sh TAG [full]
TAG=${1:-n}; FULL=$2
OUT=gpurun_out; mkdir -p $OUT
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err || { echo "bench failed"; tail -5 $OUT/bench_$TAG.err; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k 'regex:^k_' -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_launch_$TAG.log 2>&1; echo "ncu list rc=$?"
if [ -n "$FULL" ]; then
ncu --set full --clock-control none --import-source on -k 'regex:k_interp|k_near_search|k_near_slow|k_flags|k_walk|k_valid|k_mesh_weights|k_scan_counts|k_compact|k_path_fix' \
    --launch-skip 33 --launch-count 11 -o $OUT/prof_$TAG -f \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
fi
