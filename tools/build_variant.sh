#!/bin/bash
# Builds build/variants/lib_<name>.so with extra -D flags for integrator.cu (A/B experiments; select with ASUNA_B200_LIB).
# usage: tools/build_variant.sh <name> <nvcc flags...>
set -e
cd "$(dirname "$0")/../asuna_b200/csrc"
name=$1; shift
mkdir -p ../../build/variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
nvcc $ARCH -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c integrator.cu -o /tmp/integrator_$name.o
nvcc $ARCH -shared -o ../../build/variants/lib_$name.so api.o bvh_build.o /tmp/integrator_$name.o
cuobjdump -res-usage /tmp/integrator_$name.o 2>&1 | grep -A1 "k_trace_closestILb0" | grep REG
