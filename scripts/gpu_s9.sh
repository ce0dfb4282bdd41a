#!/bin/bash
OUT=gpurun_out/r01s9; mkdir -p $OUT
timeout 900 python -m pytest tests/test_solve_gpu.py -q -x -k two_level --timeout 600 > $OUT/pytest_amr.log 2>&1; echo "exit $?" >> $OUT/pytest_amr.log
FUSION=0 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_gsrb_pair" -c 2 -f -o $OUT/prof_pair python scripts/prof_kernels.py 512 smooth > $OUT/ncu_pair.log 2>&1
FUSION=1 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"k_gsrb2" -c 1 -f -o $OUT/prof_fused python scripts/prof_kernels.py 512 smooth > $OUT/ncu_fused.log 2>&1
ls -la $OUT
