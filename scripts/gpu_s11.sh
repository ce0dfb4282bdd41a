#!/bin/bash
OUT=gpurun_out/r01s11; mkdir -p $OUT
TUNE_SKIP_FUSED=1 timeout 600 python scripts/tune_smoother.py 512 10 > $OUT/tune_lean.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
