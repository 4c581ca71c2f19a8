#!/bin/bash
# round 2, visit h (2 GPUs): seam kernel (k_near_weak) + chunked push + fused finish: parity, K1 timings, N=2 with verification
TAG=${1:-r2h}
OUT=gpurun_out; mkdir -p $OUT
t0=$SECONDS
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 400 python -m pytest tests -m gpu -q -x --durations=4 --deselect tests/test_gpu_fullsize.py > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$? ($((SECONDS-t0)) s)"; tail -3 $OUT/pytest_$TAG.log
timeout 120 python scripts/prof_path.py C3 3 > $OUT/prof_C3_$TAG.log 2>&1; tail -3 $OUT/prof_C3_$TAG.log
AB_STAGES=K1,path timeout 90 python scripts/ab_opt.py pdl=1 > $OUT/ab_$TAG.log 2>&1; cat $OUT/ab_$TAG.log
timeout 300 $TR bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/bench_n2_$TAG.json 2> $OUT/bench_n2_$TAG.err; echo "bench N=2 rc=$? ($((SECONDS-t0)) s)"; grep -v "NCCL INFO" $OUT/bench_n2_$TAG.err | tail -6; grep -c "NCCL INFO" $OUT/bench_n2_$TAG.err; grep "nranks" $OUT/bench_n2_$TAG.err | head -3
timeout 300 $TR bench.py --gpus 2 --steps 5 --warmup 3 --barrier nccl --no-e2e > $OUT/bench_n2nccl_$TAG.json 2> $OUT/bench_n2nccl_$TAG.err; echo "bench N=2 (nccl barrier) rc=$? ($((SECONDS-t0)) s)"
python - <<'PY'
import json
for f in ("bench_n2_%s","bench_n2nccl_%s"):
    try:
        b=json.load(open("gpurun_out/"+f % "r2h"+".json")); print(f, b["ms_per_step"], b["details"]["ms_per_step_by_rank"], b["verified"], (b.get("e2e") or {}).get("ms_per_step"))
    except Exception as e: print(f, "ERR", e)
PY
timeout 300 $TR bench.py --gpus 2 --workload C5s --steps 2 --warmup 1 > $OUT/bench_c5s_n2_$TAG.json 2> $OUT/bench_c5s_n2_$TAG.err; echo "bench C5s N=2 rc=$? ($((SECONDS-t0)) s)"; grep "verify" $OUT/bench_c5s_n2_$TAG.err | tail -2; cut -c1-200 $OUT/bench_c5s_n2_$TAG.json
timeout 200 python bench.py --no-cpu > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err; echo "bench N=1 rc=$? ($((SECONDS-t0)) s)"; tail -2 $OUT/bench_$TAG.err
echo "done ($((SECONDS-t0)) s)"
