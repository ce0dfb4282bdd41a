#!/bin/bash
OUT=gpurun_out/r01s17; mkdir -p $OUT
timeout 900 python scripts/tune_smoother.py 512 10 > $OUT/tune_v3.log 2>&1; echo "exit $?" >> $OUT/tune_v3.log
timeout 900 python -m pytest tests/test_solve_gpu.py tests/test_prims_gpu.py -q -x --timeout 600 > $OUT/pytest.log 2>&1; echo "exit $?" >> $OUT/pytest.log
