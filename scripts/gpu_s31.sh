#!/bin/bash
OUT=gpurun_out/r01s31; mkdir -p $OUT
timeout 900 python -m pytest tests -q -m gpu -x > $OUT/pytest_gpu.log 2>&1; echo "exit $?" >> $OUT/pytest_gpu.log
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --profile-out $OUT/prof1_512.txt > $OUT/b1_512.log 2> $OUT/b1_512.err; echo "exit $?" >> $OUT/b1_512.err
B200MG_NO_FUSED_RESNORM=1 B200MG_NO_ZERO_INPUT=1 B200MG_NO_BC_OVERLAP=1 timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/b1_512_alloff.log 2> $OUT/b1_512_alloff.err; echo "exit $?" >> $OUT/b1_512_alloff.err
B200MG_NO_BC_OVERLAP=1 timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/b1_512_nobcov.log 2> $OUT/b1_512_nobcov.err; echo "exit $?" >> $OUT/b1_512_nobcov.err
