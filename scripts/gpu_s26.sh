#!/bin/bash
OUT=gpurun_out/r01s26; mkdir -p $OUT
timeout 600 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "exit $?" >> $OUT/pytest_gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --profile-out $OUT/prof2_512.txt > $OUT/b2_512.log 2> $OUT/b2_512.err; echo "exit $?" >> $OUT/b2_512.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --fusion 0 > $OUT/b2_512_nofuse.log 2> $OUT/b2_512_nofuse.err; echo "exit $?" >> $OUT/b2_512_nofuse.err
