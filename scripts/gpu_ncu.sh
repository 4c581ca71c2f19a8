#!/bin/bash
# ncu launch list (+ optional --set full captures) of the bench workload.
#   bash scripts/gpu_ncu.sh TAG [full]
TAG=${1:-n}; FULL=$2
OUT=gpurun_out; mkdir -p $OUT
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err || { echo "bench failed"; tail -5 $OUT/bench_$TAG.err; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k 'regex:^k_' -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_launch_$TAG.log 2>&1; echo "ncu list rc=$?"
if [ -n "$FULL" ]; then
# (a) the ten kernels of one gzb_map_interp step (the 4th one: after three warm-up steps)
ncu --set full --clock-control none --import-source on -k 'regex:k_near_search|k_near_slow|k_chain|k_walk|k_valid|k_mesh_weights_interp|k_compact|k_path_fix' \
    --launch-skip 30 --launch-count 10 -o $OUT/prof_$TAG -f \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full (step) rc=$?"
# (b) the stand-alone stage kernels of the per-stage table: K2+K3 fused, K2, K3, K4
ncu --set full --clock-control none --import-source on -k 'regex:^k_interp' \
    --launch-skip 1 --launch-count 1 -o $OUT/prof_${TAG}_k4 -f \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_full_${TAG}_k4.log 2>&1; echo "ncu full (K4) rc=$?"
ncu --set full --clock-control none --import-source on -k 'regex:^k_mesh_weights$' \
    --launch-skip 1 --launch-count 1 -o $OUT/prof_${TAG}_k2k3 -f \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_full_${TAG}_k2k3.log 2>&1; echo "ncu full (K2K3) rc=$?"
fi
