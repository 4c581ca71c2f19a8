#!/bin/bash
# round 2, visit y: (1 GPU) the tests touched last
OUT=gpurun_out; mkdir -p $OUT
t0=$SECONDS
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -k "launch_and_memory or c3_chain_links" > $OUT/pytest_r2y.log 2>&1; echo "pytest rc=$? ($((SECONDS-t0)) s)"; tail -3 $OUT/pytest_r2y.log
