#!/bin/bash
# Scaling legs on an N-GPU box: bash scripts/gpu_scale.sh N TAG   -> gpurun_out/scale_<workload>_n<N>_<TAG>.json
# C3 weak (the driver's default), C5 strong (6 missions x 1 year), C4 strong (ORCA36-class, 30 days); verification on.
N=${1:-2}; TAG=${2:-r2}
OUT=gpurun_out; mkdir -p $OUT
t0=$SECONDS
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"
run() {   # name, args...
  local name=$1; shift
  timeout 420 $TR bench.py --gpus $N "$@" > $OUT/scale_${name}_n${N}_$TAG.json 2> $OUT/scale_${name}_n${N}_$TAG.err
  echo "== $name N=$N rc=$? ($((SECONDS-t0)) s)"; grep -v "NCCL INFO" $OUT/scale_${name}_n${N}_$TAG.err | grep -i "verify\|error\|Traceback\|differs" | tail -3
  python - $OUT/scale_${name}_n${N}_$TAG.json <<'PY'
import json, sys
try:
    b = json.load(open(sys.argv[1]))
    e = b.get("e2e") or {}
    print("   value %.4g pts/s  ms/step %.4f  by rank %s  verified %s  e2e %s ms  reseeds %s" % (
        b["value"], b["ms_per_step"], b["details"].get("ms_per_step_by_rank"), b.get("verified"), e.get("ms_per_step"), b["details"].get("seed_reseeds")))
except Exception as exc:
    print("   no JSON line:", exc)
PY
}
run C3 --steps 10 --warmup 3
run C5 --workload C5 --steps 3 --warmup 1
run C4 --workload C4 --steps 3 --warmup 1
if [ "$N" = "8" ]; then
  run C3nccl --steps 10 --warmup 3 --barrier nccl --no-e2e
  run C3gathernccl --steps 10 --warmup 3 --gather nccl --no-e2e
fi
grep -h "nranks" $OUT/scale_C3_n${N}_$TAG.err | head -2
echo "done ($((SECONDS-t0)) s)"
