#!/bin/bash
OUT=gpurun_out/r01s35; mkdir -p $OUT
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --profile-out $OUT/prof2_512.txt > $OUT/b2_512.log 2> $OUT/b2_512.err; echo "exit $?" >> $OUT/b2_512.err
