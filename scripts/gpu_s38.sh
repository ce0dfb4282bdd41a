#!/bin/bash
OUT=gpurun_out/r01s38; mkdir -p $OUT
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "exit $?" >> $OUT/smoke.log
timeout 120 python scripts/gmres_vs_mlmg.py 256 > $OUT/gmres_256.log 2>&1; echo "exit $?" >> $OUT/gmres_256.log
timeout 120 python -m pytest tests/test_solve_gpu.py -q -m gpu -k "neumann" > $OUT/pytest_neumann.log 2>&1; echo "exit $?" >> $OUT/pytest_neumann.log
