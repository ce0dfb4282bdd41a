#!/bin/bash
OUT=gpurun_out/r01s40; mkdir -p $OUT; cd $OUT
( timeout 8 ../../build/refdriver/ABecLaplacian_C.b200.ex max_level=0 n_cell=64 max_grid_size=32 prob_type=2 verbose=2 composite_solve=1 > driver_p2.log 2>&1; echo "exit $?" >> driver_p2.log; ls plot plot/Level_0 >> driver_p2.log 2>&1 )
