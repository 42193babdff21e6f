#!/bin/bash
# Build compile-time variants of libgg_b200.so HERE (nvcc cross-compiles); the .so files travel to the GPU box with the snapshot.
#   tools/build_variants.sh name1:"-DFLAG ..." name2:"..."
# Only gg_kernels.cu is recompiled with the flags; the stacking unit is linked from the regular build (csrc/build/gg_stack.o).
cd "$(dirname "$0")/.."
mkdir -p gym_genesis_b200/variants
python -c "import __graft_entry__ as g; g.build_cuda()" || exit 1
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  ( nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $flags \
       -I include -I gym_genesis_b200/csrc -c gym_genesis_b200/csrc/gg_kernels.cu -o gym_genesis_b200/variants/$name.o && \
    nvcc -shared -gencode arch=compute_100a,code=sm_100a -o gym_genesis_b200/variants/$name.so gym_genesis_b200/variants/$name.o gym_genesis_b200/csrc/build/gg_stack.o && \
    rm -f gym_genesis_b200/variants/$name.o ) &
done
wait
for so in gym_genesis_b200/variants/*.so; do GG_B200_LIB=$PWD/$so python -c "from gym_genesis_b200 import cabi; cabi.load_library(); print('loads:', '$so')" || exit 1; done
