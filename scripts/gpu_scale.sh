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
if [ "$N" = "8" ] || [ "$N" = "4" ]; then
  run C3chunks1 --steps 10 --warmup 3 --no-e2e --bulk-chunks 1
  run C3chunks4 --steps 10 --warmup 3 --no-e2e --bulk-chunks 4
  run C3chunks2 --steps 10 --warmup 3 --no-e2e --bulk-chunks 2
fi
if [ "$N" = "8" ]; then
  run C3nccl --steps 10 --warmup 3 --barrier nccl --no-e2e
  run C3gathernccl --steps 10 --warmup 3 --gather nccl --no-e2e
fi
grep -h "nranks" $OUT/scale_C3_n${N}_$TAG.err | head -2
if [ "$N" = "8" ]; then      # the smaller worlds on the same box (what the driver's scaling run does): C3 weak at 1, 2, 4
  for M in 1 2 4; do
    TRM="python -m torch.distributed.run --nnodes=1 --nproc-per-node $M --master-addr 127.0.0.1 --master-port 29543"
    timeout 300 $TRM bench.py --gpus $M --steps 10 --warmup 3 --no-cpu --no-f2 > $OUT/scale_C3_n${M}on8_$TAG.json 2> $OUT/scale_C3_n${M}on8_$TAG.err
    python - $OUT/scale_C3_n${M}on8_$TAG.json $M <<'PY'
import json, sys
b = json.load(open(sys.argv[1])); e = b.get("e2e") or {}
print("   same box, N=%s: value %.4g pts/s  ms/step %.4f  e2e %s ms  verified %s" % (sys.argv[2], b["value"], b["ms_per_step"], e.get("ms_per_step"), b.get("verified")))
PY
  done
fi
echo "done ($((SECONDS-t0)) s)"
