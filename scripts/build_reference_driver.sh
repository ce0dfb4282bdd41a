#!/bin/bash
# Compiles the reference's own test driver Tests/LinearSolvers/ABecLaplacian_C (main.cpp, MyTest.cpp, initProb.cpp,
# MyTestPlotfile.cpp) UNMODIFIED, from where it lies, against this library's headers (amrex_b200/csrc/compat forwards the
# reference's include names) and links it with libamrex_b200.so.  nvcc flags are the ones the reference's CUDA build uses
# for application code (--extended-lambda --expt-relaxed-constexpr).  No GPU is needed to build.
# The executable uses the SAME (shared) CUDA runtime as the library: streams and events cross the boundary.
#   scripts/build_reference_driver.sh [reference_root] [out_dir]      -> <out_dir>/ABecLaplacian_C.b200.ex
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
REF=${1:-/root/reference}
OUT=${2:-$ROOT/build/refdriver}
D=$REF/Tests/LinearSolvers/ABecLaplacian_C
mkdir -p "$OUT"
INC="-I$ROOT/amrex_b200/csrc/compat -I$ROOT/amrex_b200/csrc/base -I$ROOT/amrex_b200/csrc/mlmg -I$ROOT/include -I$D"
for f in main MyTest initProb MyTestPlotfile; do
    nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 --extended-lambda --expt-relaxed-constexpr -x cu $INC -c "$D/$f.cpp" -o "$OUT/$f.o"
done
nvcc -cudart shared -gencode arch=compute_100a,code=sm_100a "$OUT"/main.o "$OUT"/MyTest.o "$OUT"/initProb.o "$OUT"/MyTestPlotfile.o -o "$OUT/ABecLaplacian_C.b200.ex" \
     -L"$ROOT/amrex_b200/lib" -lamrex_b200 -Xlinker -rpath -Xlinker "$ROOT/amrex_b200/lib"
echo "built $OUT/ABecLaplacian_C.b200.ex"
