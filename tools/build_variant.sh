#!/bin/bash
# Variant library of the tensor residual kernel: tools/build_variant.sh <name> "<extra nvcc -D flags>"
# -> pde_inverse_problem_b200/libpdeip_<name>.so (all other objects from csrc/build; run `make` there first).
# Use with PDEIP_LIB=$PWD/pde_inverse_problem_b200/libpdeip_<name>.so.  Variant libraries are git-ignored.
set -e
cd "$(dirname "$0")/../pde_inverse_problem_b200/csrc"
name=$1; shift
mkdir -p build_var
nvcc -O3 -std=c++17 -lineinfo -DPDEIP_HAVE_TENSOR_PATH -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC \
  -I../../include -I. --expt-relaxed-constexpr $@ -c residual_tensor.cu -o build_var/rt_$name.o
objs=$(ls build/*.o | grep -v residual_tensor.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libpdeip_$name.so $objs build_var/rt_$name.o -lcudart
echo built ../libpdeip_$name.so
