#!/bin/bash
OUT=gpurun_out/r01s19; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2956$N bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --profile-out $OUT/prof${N}_512.txt > $OUT/b${N}_512.log 2> $OUT/b${N}_512.err; echo "exit $?" >> $OUT/b${N}_512.err
done
