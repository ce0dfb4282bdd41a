#!/bin/bash
# Developer tool: builds libamrex_b200 with build-time variants of one kernel file for A/B timing on the GPU box
# (python scripts/tune_smoother.py under AMREX_B200_LIB=<variant library>).
#   scripts/build_variants.sh <kernel file (relative to amrex_b200/csrc/kernels)> name1:"-DX=1 -DY=0" name2:"..." ...
# Output: amrex_b200/lib/var/libamrex_b200_<name>.so (git-ignored, travels with gpurun).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$1; shift
cd "$ROOT/amrex_b200/csrc"
make -j8 > /dev/null
mkdir -p ../lib/var ../../build/var
BASE=$(basename "$SRC" .cu)
OBJS=$(ls ../../build/obj/kernels/*.o ../../build/obj/base/*.o ../../build/obj/mlmg/*.o ../../build/obj/compat/*.o ../../build/obj/capi/*.o | grep -v "/$BASE.o")
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  (
    src=kernels/$SRC
    if [ -f "$flags" ]; then src=$flags; flags=""; fi        # name:/path/to/other_source.cu
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 --extended-lambda \
         -Xcompiler -fPIC,-Wall,-Wno-unused-function -I../../include -Ikernels $flags -c $src -o ../../build/var/${BASE}_$name.o
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/var/libamrex_b200_$name.so ../../build/var/${BASE}_$name.o $OBJS -lcudart -lnccl
    echo "built $name"
  ) &
done
wait
