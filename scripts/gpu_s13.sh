#!/bin/bash
OUT=gpurun_out/r01s13; mkdir -p $OUT
run2() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e "${@:2}"; }
timeout 300 bash -c "$(declare -f run2); run2 29521 --n-cell 128 --max-grid-size 32" > $OUT/b_128_32.log 2> $OUT/b_128_32.err; echo "exit $?" >> $OUT/b_128_32.err
timeout 300 bash -c "$(declare -f run2); run2 29522 --n-cell 256 --max-grid-size 64" > $OUT/b_256_64.log 2> $OUT/b_256_64.err; echo "exit $?" >> $OUT/b_256_64.err
# memcheck of the smallest failing configuration
if grep -q "exit 1" $OUT/b_128_32.err; then CFG="--n-cell 128 --max-grid-size 32"; elif grep -q "exit 1" $OUT/b_256_64.err; then CFG="--n-cell 256 --max-grid-size 64"; else CFG="--n-cell 512 --max-grid-size 128"; fi
echo "memcheck config: $CFG" > $OUT/memcheck.log
timeout 1200 compute-sanitizer --tool memcheck --target-processes all --print-limit 20 --log-file $OUT/sanitizer_%p.log python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e $CFG >> $OUT/memcheck.log 2>&1
for f in $OUT/sanitizer_*.log; do echo "== $f"; head -60 $f; done > $OUT/sanitizer_summary.txt 2>&1
rm -f $OUT/sanitizer_*.log
