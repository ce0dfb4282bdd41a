#!/bin/bash
OUT=gpurun_out/r01s18; mkdir -p $OUT
timeout 600 python scripts/debug_fused3.py > $OUT/debug_fused3.log 2>&1; echo "exit $?" >> $OUT/debug_fused3.log
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
