#!/bin/bash
OUT=gpurun_out/r01s30; mkdir -p $OUT
TUNE_PLANS="4,8,4,2,0;4,6,4,2,0;4,4,4,4,0" timeout 300 python scripts/tune_fused4.py 256 20 64 > $OUT/tune4_256_mgs64.log 2>&1; echo "exit $?" >> $OUT/tune4_256_mgs64.log
timeout 900 python -m pytest tests -q -m gpu -x > $OUT/pytest_gpu.log 2>&1; echo "exit $?" >> $OUT/pytest_gpu.log
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --profile-out $OUT/prof1_512.txt > $OUT/b1_512.log 2> $OUT/b1_512.err; echo "exit $?" >> $OUT/b1_512.err
