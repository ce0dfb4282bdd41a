#!/bin/bash
OUT=gpurun_out/r01s14; mkdir -p $OUT
B200MG_DEBUG_SYNC=1 NCCL_DEBUG=WARN timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --n-cell 128 --max-grid-size 32 > $OUT/dbg.log 2> $OUT/dbg.err; echo "exit $?" >> $OUT/dbg.err
