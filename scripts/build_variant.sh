#!/bin/bash
# Development aid: build the C-ABI library with extra compile-time switches into variants/lib_<name>.so for A/B
# timing on the GPU box (scripts/f0_probe.py and scripts/inv_probe.py honour RIC_LIB=<path>).
# usage: scripts/build_variant.sh <name> [extra nvcc flags, e.g. -DRIC_EXP_FLATQ=0]
set -e
cd "$(dirname "$0")/../rududu_image_codec_b200/csrc"
name=$1; shift
mkdir -p ../../variants
make -s ric_entropy.o ric_entropy_gpu.o
nvcc --threads 4 -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-O2 -Xptxas -v "$@" \
     -c -o /tmp/ric_b200_$name.o ric_b200.cu 2> ../../variants/build_$name.log
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../variants/lib_$name.so /tmp/ric_b200_$name.o ric_entropy_gpu.o ric_entropy.o -lcudart -lpthread
grep -A3 "fwd_level_kernelILb1ELi0ELi1E\|inv_level_kernelILb1ELi0ELi2E" ../../variants/build_$name.log | grep "Used" || true
