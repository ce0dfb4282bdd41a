#!/bin/bash
OUT=gpurun_out/r01s22; mkdir -p $OUT
timeout 300 python scripts/debug_fused4.py > $OUT/debug_fused4.log 2>&1; echo "exit $?" >> $OUT/debug_fused4.log
timeout 300 python scripts/tune_fused4.py 512 10 > $OUT/tune4_512.log 2>&1; echo "exit $?" >> $OUT/tune4_512.log
TUNE_PLANS="4,8,4,3" timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gsrb4 -s 3 -c 1 -o $OUT/gsrb4_843 python scripts/tune_fused4.py 512 2 > $OUT/ncu.log 2>&1; echo "exit $?" >> $OUT/ncu.log
