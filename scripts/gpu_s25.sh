#!/bin/bash
OUT=gpurun_out/r01s25; mkdir -p $OUT
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --profile-out $OUT/prof1_512.txt > $OUT/b1_512.log 2> $OUT/b1_512.err; echo "exit $?" >> $OUT/b1_512.err
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --fusion 0 --no-cpu-baseline --no-e2e > $OUT/b1_512_nofuse.log 2> $OUT/b1_512_nofuse.err; echo "exit $?" >> $OUT/b1_512_nofuse.err
