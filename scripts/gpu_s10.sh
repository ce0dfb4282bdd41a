#!/bin/bash
OUT=gpurun_out/r01s10; mkdir -p $OUT
timeout 600 python scripts/debug_amr.py 1 32 16 3 > $OUT/debug_amr_p1.log 2>&1
timeout 600 python scripts/debug_amr.py 2 32 16 3 > $OUT/debug_amr_p2.log 2>&1
