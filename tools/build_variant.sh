#!/bin/bash
# dev tool: build a Panda-only library variant into variants/lib_<name>.so   usage: build_variant.sh name [nvcc flags...]
set -e
name=$1; shift
cd "$(dirname "$0")/../vamp_mvt_b200/csrc"
mkdir -p ../../variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -DVMV_DEV_PANDA_ONLY -split-compile 0 "$@" -o ../../variants/lib_$name.so vmv_host.cu 2>&1 | grep -E "error|registers.*v4" | head
echo built $name
