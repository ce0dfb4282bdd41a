#!/bin/bash
OUT=gpurun_out/r01s36; mkdir -p $OUT
# 1. the judged bench line (N=1, e2e + CPU baseline)
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --profile-out $OUT/prof1_512.txt > $OUT/b1_512.log 2> $OUT/b1_512.err; echo "exit $?" >> $OUT/b1_512.err
# 2. launch list of the same command (shorter run), ncu per-launch durations
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches.csv python bench.py --gpus 1 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/ncu_list.log 2>&1; echo "exit $?" >> $OUT/ncu_list.log
# 3. full capture of the dominant kernel (one launch, after warm-up) + its zero-input variant
TUNE_PLANS="4,8,4,2,0" timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gsrb4 -s 3 -c 1 -o $OUT/gsrb4_842 python scripts/tune_fused4.py 512 2 > $OUT/ncu_full.log 2>&1; echo "exit $?" >> $OUT/ncu_full.log
# 4. the reference arm, as the driver runs it
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > $OUT/ref_arm.log 2> $OUT/ref_arm.err; echo "exit $?" >> $OUT/ref_arm.err
