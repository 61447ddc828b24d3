#!/bin/bash
# build a tuning variant of the library: scripts/build_variant.sh <name> <extra nvcc flags...>
set -e
name=$1; shift
cd "$(dirname "$0")/../bayeslim_b200/csrc"
mkdir -p variants
objs=""
for f in fringe_kernels builder_kernels capi; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v "$@" -c $f.cu -o variants/${name}_$f.o 2> variants/${name}_$f.log
  objs="$objs variants/${name}_$f.o"
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/lib_${name}.so $objs
rm -f $objs
grep -A2 "Lb1" variants/${name}_fringe_kernels.log | grep -E "Used|spill" | grep -B1 "Used" | paste - - | sed 's/ptxas info    : //' | grep "If" -A0 | head -0
grep -E "Compiling.*I[f]Lb1|Used|spill" variants/${name}_fringe_kernels.log | grep -A2 "IfLb1" | grep -E "Used|spill" | paste - - | cut -c1-150
