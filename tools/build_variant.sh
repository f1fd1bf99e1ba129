#!/bin/bash
# build a variant of the native library with extra nvcc defines:  tools/build_variant.sh <out.so> [-DNAME=VAL ...]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$1; shift
CS=$ROOT/amatsukaze_b200/csrc
mkdir -p "$(dirname "$OUT")"
g++ -std=c++17 -O2 -fPIC -ffp-contract=off -fno-fast-math -fvisibility=hidden -Wno-unknown-pragmas -c $CS/logo_host.cpp -o /tmp/logo_host_var.o
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
  --expt-relaxed-constexpr --extended-lambda -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math,-fvisibility=hidden "$@" \
  -shared -o "$OUT" $CS/amtk_b200.cu /tmp/logo_host_var.o -ldl -lpthread
