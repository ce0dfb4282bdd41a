#!/bin/bash
OUT=gpurun_out/r01s15; mkdir -p $OUT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --n-cell 128 --max-grid-size 32 > $OUT/b2_128.log 2> $OUT/b2_128.err; echo "exit $?" >> $OUT/b2_128.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --profile-out $OUT/prof2_512.txt > $OUT/b2_512.log 2> $OUT/b2_512.err; echo "exit $?" >> $OUT/b2_512.err
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --profile-out $OUT/prof1_512.txt > $OUT/b1_512.log 2> $OUT/b1_512.err; echo "exit $?" >> $OUT/b1_512.err
