#!/bin/bash
# Build-time variants of libabm_b200.so for one screening round trip on the GPU box (variants/*.so: git-ignored, they
# travel with the tree).   usage: tools/build_variants.sh name "nvcc -D flags" [name "flags" ...]
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared $flags \
      -I/usr/include -o variants/$name.so covid19-agent-based-model_b200/csrc/abm_capi.cu -ldl
  echo "built variants/$name.so ($flags)"
done
