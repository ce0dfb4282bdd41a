#!/bin/bash
OUT=gpurun_out/r01s32; mkdir -p $OUT
timeout 900 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "exit $?" >> $OUT/pytest_gpu.log
TUNE_PLANS="4,8,4,2,0" timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_gsrb_shell|k_copy_tags|k_apply_bc" -s 24 -c 5 -o $OUT/o2_kernels python scripts/tune_fused4.py 512 2 > $OUT/ncu.log 2>&1; echo "exit $?" >> $OUT/ncu.log
