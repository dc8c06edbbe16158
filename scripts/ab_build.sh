#!/bin/bash
# Build a variant of the library from the CURRENT working tree into ab_libs/lib<name>.so (own source copy, so
# several variants can compile in parallel):  scripts/ab_build.sh <name> &
set -e
cd "$(dirname "$0")/.."
name=$1
mkdir -p ab_src/$name ab_libs
rm -rf ab_src/$name/*
cp -r mcmcdiagnostictools.jl_b200/csrc ab_src/$name/csrc
mkdir -p ab_src/$name/include && cp include/mcmcdiag_b200.h ab_src/$name/include/
sed -i 's|#include "../../include/mcmcdiag_b200.h"|#include "../include/mcmcdiag_b200.h"|' ab_src/$name/csrc/mcd_api.cu
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared \
  -o ab_libs/lib$name.so ab_src/$name/csrc/mcd_api.cu > ab_src/$name/build.log 2>&1
echo "built $name"
