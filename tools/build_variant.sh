#!/bin/bash
# usage: tools/build_variant.sh <name> [-DFLAG=1 ...]  -> flowchain_iccv2023_b200/libdev_<name>.so (git-ignored; travels with gpurun)
# run a bench against it with FLOWCHAIN_B200_LIB=flowchain_iccv2023_b200/libdev_<name>.so
name=$1; shift
cd "$(dirname "$0")/../flowchain_iccv2023_b200/csrc" || exit 1
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC "$@" -o ../libdev_$name.so capi.cu 2>&1 | grep -E "error" ; ls -la ../libdev_$name.so
