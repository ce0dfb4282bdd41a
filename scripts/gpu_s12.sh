#!/bin/bash
OUT=gpurun_out/r01s12; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 --n-cell 256 --no-cpu-baseline > $OUT/bench2_256.log 2> $OUT/bench2_256.err; echo "exit $?" >> $OUT/bench2_256.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench2_512.log 2> $OUT/bench2_512.err; echo "exit $?" >> $OUT/bench2_512.err
timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > $OUT/bench1_512.log 2> $OUT/bench1_512.err; echo "exit $?" >> $OUT/bench1_512.err
