#!/bin/bash
# Builds build/variants/lib_<name>.so with extra -D flags for bvh_build.cu (builder A/B experiments; select with ASUNA_B200_LIB).
# usage: tools/build_variant_bvh.sh <name> <nvcc flags...>
set -e
cd "$(dirname "$0")/../asuna_b200/csrc"
name=$1; shift
mkdir -p ../../build/variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
nvcc $ARCH -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c bvh_build.cu -o /tmp/bvh_build_$name.o
nvcc $ARCH -shared -o ../../build/variants/lib_$name.so api.o integrator.o /tmp/bvh_build_$name.o
cuobjdump -res-usage /tmp/bvh_build_$name.o 2>&1 | grep -A1 "k_ploc\|k_emit_wide" | grep -o "REG:[0-9]*" | tr '\n' ' '; echo
