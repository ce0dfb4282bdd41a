#!/bin/bash
OUT=gpurun_out/r01s34; mkdir -p $OUT
timeout 600 python -m pytest tests/test_solve_gpu.py -q -m gpu -k "fcycle or neumann or fluxes" > $OUT/pytest_gpu.log 2>&1; echo "exit $?" >> $OUT/pytest_gpu.log
