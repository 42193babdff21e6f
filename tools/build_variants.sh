#!/bin/bash
# Build compile-time variants of libgg_b200.so HERE (nvcc cross-compiles); the .so files travel to the GPU box with the snapshot.
#   tools/build_variants.sh name1:"-DFLAG ..." name2:"..."
cd "$(dirname "$0")/.."
mkdir -p gym_genesis_b200/variants
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared $flags \
       -I include -I gym_genesis_b200/csrc -o gym_genesis_b200/variants/$name.so gym_genesis_b200/csrc/gg_kernels.cu &
done
wait
ls -la gym_genesis_b200/variants
