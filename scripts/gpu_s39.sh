#!/bin/bash
OUT=gpurun_out/r01s39; mkdir -p $OUT
export B200MG_BOTTOM_KERNEL=1
timeout 150 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "exit $?" >> $OUT/pytest_gpu.log
timeout 60 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --profile-out $OUT/prof1_512.txt > $OUT/b1_512.log 2> $OUT/b1_512.err; echo "exit $?" >> $OUT/b1_512.err
