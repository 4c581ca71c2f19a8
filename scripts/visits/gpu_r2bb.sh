#!/bin/bash
# round 2, visit bb (1 GPU): medium-size pageable uploads through the copy threads when the stream is idle
OUT=gpurun_out; mkdir -p $OUT
t0=$SECONDS
timeout 300 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_edge.py tests/test_gpu_parity.py -m gpu -q -x > $OUT/pytest_r2bb.log 2>&1; echo "pytest rc=$? ($((SECONDS-t0)) s)"; tail -2 $OUT/pytest_r2bb.log
timeout 100 python scripts/e2e_phases.py 3 > $OUT/e2e_phases_r2bb.log 2>&1; grep "rep 2" $OUT/e2e_phases_r2bb.log | cut -c1-700; grep "Model2SatTrack" $OUT/e2e_phases_r2bb.log | tail -2
echo "done ($((SECONDS-t0)) s)"
