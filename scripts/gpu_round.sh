#!/bin/bash
# One GPU-box session: parity tests, benchmark, ncu launch list and a full capture of the dominant kernel.
# Usage (from the repo root, under gpurun):  bash scripts/gpu_round.sh [tag] [what...]   what: tests bench ncu_list ncu_full
TAG=${1:-r01}; shift
WHAT=${@:-tests bench ncu_list ncu_full}
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $OUT/gpu.txt 2>&1
nproc > $OUT/nproc.txt
for w in $WHAT; do
case $w in
tests)
  timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log ;;
smoke)
  timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log ;;
bench)
  timeout 1200 python bench.py --steps 5 --warmup 3 > $OUT/bench.log 2> $OUT/bench.err; echo "bench exit $?" >> $OUT/bench.err ;;
bench256)
  timeout 600 python bench.py --steps 5 --warmup 3 --n-cell 256 --no-cpu-baseline > $OUT/bench256.log 2> $OUT/bench256.err ;;
bench_ref)
  timeout 1200 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.log 2> $OUT/bench_ref.err ;;
ncu_list)
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches.csv \
     python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > $OUT/ncu_list.log 2>&1 ;;
ncu_full)
  # the dominant kernel (fused red+black pass, generation 5) on the finest level of the benchmark workload
  REPS=2 timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_gsrb5 -c 2 -f -o $OUT/k_gsrb5 \
     python scripts/prof_kernels.py 512 smooth > $OUT/ncu_full.log 2>&1 ;;
ncu_k)
  # full capture of the finest-level launches of every hot kernel of the path
  timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"${KREGEX:-k_gsrb5|k_gsrb4|k_adotx|k_gsrb_shell|k_copy_tags|k_apply_bc}" -c ${KCOUNT:-9} \
     -f -o $OUT/prof_kernels python scripts/prof_kernels.py ${NCELL:-512} smooth residual > $OUT/ncu_k.log 2>&1
  ls -la $OUT/prof_kernels.ncu-rep ;;
tune)
  # fused-pass generations side by side (4th field: 2 = generation 5, 1 = generation 4), same run, same box
  timeout 900 python scripts/tune_smoother.py ${NCELL:-512} 10 8,4,2,2 8,4,2,1 > $OUT/tune.log 2>&1; echo "tune exit $?" >> $OUT/tune.log ;;
bench_f0)
  timeout 900 python bench.py --steps 5 --warmup 3 --fusion 0 --no-cpu-baseline > $OUT/bench_f0.log 2> $OUT/bench_f0.err; echo "exit $?" >> $OUT/bench_f0.err ;;
bench_f1)
  timeout 900 python bench.py --steps 5 --warmup 3 --fusion 1 --no-cpu-baseline > $OUT/bench_f1.log 2> $OUT/bench_f1.err; echo "exit $?" >> $OUT/bench_f1.err ;;
esac
done
ls -la $OUT
