#!/bin/bash
# Time every variant library (tools/build_variants.sh) with the developer micro-benchmarks.  Output -> gpurun_out/variants_<tag>.log
TAG=${1:-var}
mkdir -p gpurun_out
for so in gym_genesis_b200/variants/*.so; do
  name=$(basename $so .so)
  GG_B200_LIB=$PWD/$so python tools/quick_bench.py --tag $name 2>&1 | tail -1 | tee -a gpurun_out/variants_${TAG}.log
  if [ -n "$GRASP" ]; then GG_B200_LIB=$PWD/$so python tools/grasp_bench.py 2>&1 | tail -1 | tee -a gpurun_out/variants_${TAG}.log; fi
done
