#!/bin/bash
# Build a variant of the library from the CURRENT working tree into ab_libs/lib<name>.so:
#   scripts/ab_build.sh <name> [extra nvcc flags, e.g. -DRK_MERGE_LOOKAHEAD=0]
# Only the headline kernel's translation unit (mcd_rk2.cu) is recompiled (from a private copy of csrc/, so several
# variants can compile in parallel); the other objects come from the regular in-tree build (python build.py).
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p ab_src/$name ab_libs
rm -rf ab_src/$name/*
cp -r mcmcdiagnostictools.jl_b200/csrc ab_src/$name/csrc
python mcmcdiagnostictools.jl_b200/build.py > /dev/null
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" \
  -c -o ab_src/$name/mcd_rk2.o ab_src/$name/csrc/mcd_rk2.cu > ab_src/$name/build.log 2>&1
/usr/local/cuda/bin/nvcc -shared -o ab_libs/lib$name.so ab_src/$name/mcd_rk2.o \
  mcmcdiagnostictools.jl_b200/_obj/mcd_api.o mcmcdiagnostictools.jl_b200/_obj/mcd_big.o
echo "built $name"
