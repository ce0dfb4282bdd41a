#!/bin/bash
OUT=gpurun_out/r01s24; mkdir -p $OUT
timeout 300 python scripts/debug_fused4.py 3 > $OUT/debug_fused4.log 2>&1; echo "exit $?" >> $OUT/debug_fused4.log
TUNE_PLANS="4,8,4,3;4,8,4,2;4,6,5,3;4,6,4,2" timeout 300 python scripts/tune_fused4.py 512 10 > $OUT/tune4_512.log 2>&1; echo "exit $?" >> $OUT/tune4_512.log
timeout 600 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "exit $?" >> $OUT/pytest_gpu.log
