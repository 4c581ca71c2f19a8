#!/bin/bash
# N-GPU A/B of the gather options inside ONE bench process (bench.py GZB_AB): bash scripts/gpu_ab_multi.sh N TAG "settings" [bench args...]
N=${1:-2}; TAG=${2:-abm}; AB=${3:-"push_blocks=296;push_mode=1;push_mode=0,push_blocks=64,bulk_chunks=1;bulk_chunks=4"}
shift 3
OUT=gpurun_out; mkdir -p $OUT
t0=$SECONDS
GZB_AB="$AB" GZB_AB_TRACE=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29547 \
    bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --no-cpu --no-f2 "$@" > $OUT/abm_n${N}_$TAG.json 2> $OUT/abm_n${N}_$TAG.err
echo "rc=$? ($((SECONDS-t0)) s)"
grep -v "NCCL INFO" $OUT/abm_n${N}_$TAG.err | grep -o "\[rank 0\] A/B [^[]*"
grep -v "NCCL INFO" $OUT/abm_n${N}_$TAG.err | grep "gzb trace\|verify\|Error\|error\|Traceback" | head -120
python - <<PY
import json
b=json.load(open("$OUT/abm_n${N}_$TAG.json")); print("default: ms/step %.4f value %.4g verified %s by rank %s" % (b["ms_per_step"], b["value"], b.get("verified"), b["details"].get("ms_per_step_by_rank")))
PY
echo "done ($((SECONDS-t0)) s)"
