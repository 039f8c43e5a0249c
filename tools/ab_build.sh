#!/bin/bash
# Build a variant of the library for A/B runs:  tools/ab_build.sh NAME [-DFLAG ...]   ->  ab/lib_NAME.so
# (only score_tc.cu is recompiled with the extra flags; the other objects come from ocm-vae-simca_b200/build/)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p ab
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC "$@" \
     -c ocm-vae-simca_b200/csrc/score_tc.cu -o ab/score_tc_$name.o
objs=$(ls ocm-vae-simca_b200/build/*.o | grep -v score_tc.o)
nvcc -shared -o ab/lib_$name.so ab/score_tc_$name.o $objs -lcuda
cuobjdump -res-usage ab/score_tc_$name.o | grep -A1 "score_tc_kernelILb0" | grep -o "REG:[0-9]*\|STACK:[0-9]*" | paste - -
