#!/bin/bash
OUT=gpurun_out/r01s21; mkdir -p $OUT
TUNE_PLANS="4,8,4,3" timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_gsrb4 -s 3 -c 1 -o $OUT/gsrb4_843 python scripts/tune_fused4.py 512 2 > $OUT/ncu.log 2>&1; echo "exit $?" >> $OUT/ncu.log
