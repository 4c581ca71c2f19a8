#!/bin/bash
# round 2, visit v (1 GPU): delayed release of the bulk kernel (chain_first = microseconds)
TAG=${1:-r2v}
OUT=gpurun_out; mkdir -p $OUT
t0=$SECONDS
AB_STAGES=path timeout 300 python scripts/ab_opt.py chain_first=0 chain_first=4 chain_first=8 chain_first=12 chain_first=16 chain_first=24 chain_first=0 chain_first=8 chain_first=12 chain_first=16 chain_first=0 > $OUT/ab_$TAG.log 2>&1; echo "ab rc=$? ($((SECONDS-t0)) s)"; cat $OUT/ab_$TAG.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "launch_and_memory or map_interp or overlap" > $OUT/pytest_$TAG.log 2>&1; echo "pytest rc=$? ($((SECONDS-t0)) s)"; tail -2 $OUT/pytest_$TAG.log
echo "done ($((SECONDS-t0)) s)"
