#!/bin/bash
# One GPU-box visit: parity tests, the bench line, the ncu launch list and one --set full capture
# of every hot kernel.  Usage (from the repo root, on the GPU box):  bash scripts/gpu_round.sh TAG
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_$TAG.log
tail -3 $OUT/pytest_$TAG.log
python bench.py --steps 5 --warmup 3 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench rc=$?"
python bench.py --impl reference --steps 1 --warmup 0 > $OUT/bench_ref_$TAG.json 2> $OUT/bench_ref_$TAG.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_launch_$TAG.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k 'regex:k_interp|k_near_search|k_near_slow|k_flags|k_walk|k_valid|k_mesh_weights' \
    --launch-skip 24 --launch-count 8 -o $OUT/prof_$TAG -f \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > $OUT/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
cat $OUT/bench_$TAG.json
