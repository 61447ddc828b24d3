#!/bin/bash
# build a tuning variant of the library: scripts/build_tc_variant.sh <name> <extra nvcc flags...>
set -e
name=$1; shift
cd "$(dirname "$0")/../bayeslim_b200/csrc"
mkdir -p variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v "$@" -c tc_kernels.cu -o variants/${name}_tc.o 2> variants/${name}_tc.log
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/lib_${name}.so fringe_kernels.o builder_kernels.o ant_kernels.o alm_kernels.o cal_kernels.o geom_kernels.o capi.o variants/${name}_tc.o
rm -f variants/${name}_tc.o
grep -E "spill" variants/${name}_tc.log | tr '\n' ' '; echo
