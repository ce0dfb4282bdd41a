#!/bin/bash
OUT=gpurun_out/r01s27; mkdir -p $OUT
timeout 600 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "exit $?" >> $OUT/pytest_gpu.log
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/b1_512_graphs.log 2> $OUT/b1_512_graphs.err; echo "exit $?" >> $OUT/b1_512_graphs.err
B200MG_NO_GRAPHS=1 timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/b1_512_nographs.log 2> $OUT/b1_512_nographs.err; echo "exit $?" >> $OUT/b1_512_nographs.err
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --n-cell 256 > $OUT/b1_256_graphs.log 2> $OUT/b1_256_graphs.err; echo "exit $?" >> $OUT/b1_256_graphs.err
B200MG_NO_GRAPHS=1 timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --n-cell 256 > $OUT/b1_256_nographs.log 2> $OUT/b1_256_nographs.err; echo "exit $?" >> $OUT/b1_256_nographs.err
