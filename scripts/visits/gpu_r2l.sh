#!/bin/bash
# visit l (2 GPUs): how many chunks of the bulk kernel for the peer copies?
OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
for C in 1 2 4 1 2; do
  timeout 200 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --bulk-chunks $C > $OUT/chunks${C}_r2l.json 2> $OUT/chunks${C}_r2l.err
  python - $OUT/chunks${C}_r2l.json $C <<'PY'
import json, sys
b = json.load(open(sys.argv[1])); print("chunks", sys.argv[2], "ms/step %.4f" % b["ms_per_step"], b["details"]["ms_each_timed_step_rank0"][:6])
PY
done
GZB_TRACE_STEP=1 timeout 200 $TR bench.py --gpus 2 --steps 2 --warmup 3 --no-e2e --bulk-chunks 1 > /dev/null 2> $OUT/trace_n2_r2l.err; grep "gzb trace" $OUT/trace_n2_r2l.err | head -20
