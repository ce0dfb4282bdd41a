#!/bin/bash
OUT=gpurun_out/r01s28; mkdir -p $OUT
timeout 240 python scripts/debug_fused4.py 3 > $OUT/debug_fused4.log 2>&1; echo "exit $?" >> $OUT/debug_fused4.log
timeout 240 python scripts/tune_fused4.py 512 10 > $OUT/tune4_512.log 2>&1; echo "exit $?" >> $OUT/tune4_512.log
