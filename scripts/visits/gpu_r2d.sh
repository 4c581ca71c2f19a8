#!/bin/bash
# round 2, visit d (2 GPUs): multi-GPU correctness -- sharded public call, C3 weak with verification, C5s strong; + K1 timings
TAG=${1:-r2d}
OUT=gpurun_out; mkdir -p $OUT
t0=$SECONDS
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "several_chains or nearest_paths or seam or global" > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$? ($((SECONDS-t0)) s)"; tail -3 $OUT/pytest_$TAG.log
AB_STAGES=K1,path timeout 90 python scripts/ab_opt.py pdl=1 > $OUT/ab_$TAG.log 2>&1; cat $OUT/ab_$TAG.log
timeout 300 $TR bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/bench_n2_$TAG.json 2> $OUT/bench_n2_$TAG.err; echo "bench N=2 rc=$? ($((SECONDS-t0)) s)"; grep -v "NCCL INFO" $OUT/bench_n2_$TAG.err | tail -8; grep -c "NCCL INFO" $OUT/bench_n2_$TAG.err
timeout 300 $TR bench.py --gpus 2 --workload C5s --steps 2 --warmup 1 > $OUT/bench_c5s_n2_$TAG.json 2> $OUT/bench_c5s_n2_$TAG.err; echo "bench C5s N=2 rc=$? ($((SECONDS-t0)) s)"; grep -v "NCCL INFO" $OUT/bench_c5s_n2_$TAG.err | tail -6; cut -c1-250 $OUT/bench_c5s_n2_$TAG.json
timeout 200 python bench.py --workload C5s --steps 2 --warmup 1 > $OUT/bench_c5s_n1_$TAG.json 2> $OUT/bench_c5s_n1_$TAG.err; echo "bench C5s N=1 rc=$? ($((SECONDS-t0)) s)"; cut -c1-250 $OUT/bench_c5s_n1_$TAG.json
echo "done ($((SECONDS-t0)) s)"
