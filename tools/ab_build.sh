#!/bin/bash
# Build the csrc of an older commit into ab/libpangu_<tag>.so (same C ABI) for same-box A/B timing:
#   tools/ab_build.sh <commit> <tag>;  PANGU_B200_LIB=$PWD/ab/libpangu_<tag>.so python bench.py ...
# New entry points that the old sources lack are resolved lazily by ctypes only when called, but abi.lib() checks every
# prototype up front: the old build therefore also compiles the CURRENT bwd_kernels.cu / abi.cu when they only ADD symbols.
set -e
C=$1; T=$2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
W=$(mktemp -d)
git -C "$ROOT" archive "$C" pangu-pytorch-demo_b200/pangu_b200/csrc include | tar -x -C "$W"
for f in $ROOT/pangu-pytorch-demo_b200/pangu_b200/csrc/{bwd_kernels.cu,abi.cu,common.cuh}; do cp "$f" "$W/pangu-pytorch-demo_b200/pangu_b200/csrc/"; done
cp "$ROOT/include/pangu_b200.h" "$W/include/"
mkdir -p "$ROOT/ab" "$W/obj"
cd "$W/pangu-pytorch-demo_b200/pangu_b200/csrc"
for s in *.cu; do
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -I "$W/include" -I . -c "$s" -o "$W/obj/${s%.cu}.o" &
done
wait
nvcc -shared -o "$ROOT/ab/libpangu_$T.so" "$W"/obj/*.o -gencode arch=compute_100a,code=sm_100a -lcudart
ls -la "$ROOT/ab/libpangu_$T.so"
rm -rf "$W"
