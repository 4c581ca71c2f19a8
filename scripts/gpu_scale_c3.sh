#!/bin/bash
# C3 weak-scaling leg at N GPUs exactly as the driver launches it: bash scripts/gpu_scale_c3.sh N TAG
N=${1:-2}; TAG=${2:-r2f}
OUT=gpurun_out; mkdir -p $OUT
t0=$SECONDS
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 \
    > $OUT/scale_C3_n${N}_$TAG.json 2> $OUT/scale_C3_n${N}_$TAG.err
echo "rc=$? ($((SECONDS-t0)) s)"; grep -v "NCCL INFO" $OUT/scale_C3_n${N}_$TAG.err | grep -i "verify\|error\|Traceback\|e2e calls" | cut -c1-300 | tail -4
grep -h "nranks" $OUT/scale_C3_n${N}_$TAG.err | head -1 | cut -c1-200
python - <<PY
import json
b=json.load(open("$OUT/scale_C3_n${N}_$TAG.json")); e=b.get("e2e") or {}
print("value %.4g pts/s  ms/step %.4f  by rank %s  verified %s  e2e %s ms (%s pts/s)  breakdown %s" % (b["value"], b["ms_per_step"], b["details"].get("ms_per_step_by_rank"), b.get("verified"), e.get("ms_per_step"), e.get("value"), e.get("breakdown_s")))
PY
