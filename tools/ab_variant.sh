#!/bin/bash
# Build a VARIANT of the current library for same-box A/B timing: one source re-compiled with extra -D flags, linked with the
# other objects of the in-tree build (python pangu-pytorch-demo_b200/pangu_b200/build.py first):
#   tools/ab_variant.sh <tag> <source.cu> -DNAME=VALUE ...   ->  ab/libpangu_<tag>.so   (PANGU_B200_LIB=$PWD/ab/libpangu_<tag>.so)
# REPLACE=<object name> when <source.cu> is a scratch copy (e.g. an older revision) of another source.
set -e
T=$1; S=$2; shift 2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
P=$ROOT/pangu-pytorch-demo_b200/pangu_b200
W=$(mktemp -d)
mkdir -p "$ROOT/ab"
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr "$@" \
     -I "$ROOT/include" -I "$P/csrc" -c "$P/csrc/$S" -o "$W/variant.o"
R=${REPLACE:-${S%.cu}}                      # object of the in-tree build that the variant replaces (default: same name)
OBJS=$(ls "$P"/build/*.o | grep -v "/$R.o")
nvcc -shared -o "$ROOT/ab/libpangu_$T.so" $OBJS "$W/variant.o" -gencode arch=compute_100a,code=sm_100a -lcudart
rm -rf "$W"
ls -la "$ROOT/ab/libpangu_$T.so"
