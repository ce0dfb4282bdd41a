#!/bin/bash
OUT=gpurun_out/r01s20; mkdir -p $OUT
timeout 420 python scripts/debug_fused4.py > $OUT/debug_fused4.log 2>&1; echo "exit $?" >> $OUT/debug_fused4.log
timeout 300 python scripts/tune_fused4.py 512 10 > $OUT/tune4_512.log 2>&1; echo "exit $?" >> $OUT/tune4_512.log
timeout 200 python scripts/tune_fused4.py 256 10 > $OUT/tune4_256.log 2>&1; echo "exit $?" >> $OUT/tune4_256.log
