#!/bin/bash
OUT=gpurun_out/r01s37; mkdir -p $OUT
timeout 300 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "exit $?" >> $OUT/pytest_gpu.log
