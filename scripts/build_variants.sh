#!/bin/bash
# builds tuning variants of libdks.so: scripts/build_variants.sh name "-DFLAG=1 ..." [name flags ...]
mkdir -p build/variants
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC $flags \
     -Iinclude -Idistributedkernelshap_b200/csrc distributedkernelshap_b200/csrc/dks.cu -o build/variants/libdks_$name.so || exit 1
  echo built $name "($flags)"
done
