#!/bin/bash
# usage: tools/build_variant.sh <name> [extra nvcc flags...]   -> build_variants/<name>/libwbc_b200.so (hot instantiation only)
set -e
name=$1; shift
mkdir -p build_variants/$name
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC -DWBC_ONLY_HOT=1 "$@" \
  -o build_variants/$name/libwbc_b200.so mech5845m-wbc-for-legged-manipulator_b200/csrc/wbc_kernels.cu 2>&1 | grep -v "warning\|Remark\|^$\|\^\|detected during\|instantiation of\|declared but\|not reachable" | head -20
ls -la build_variants/$name/libwbc_b200.so
