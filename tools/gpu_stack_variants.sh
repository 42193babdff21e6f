#!/bin/bash
# CubeStack solver variants (compile-time builds of the SAME sources under gym_genesis_b200/variants/) + the graph-capture test
TAG=${1:-sv}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_graph.py -x -q 2>&1 | tail -5 | tee gpurun_out/${TAG}_pytest_graph.log
for v in default $(ls gym_genesis_b200/variants/ | sed 's/\.so$//'); do
  if [ $v = default ]; then unset GG_B200_LIB; else export GG_B200_LIB=$PWD/gym_genesis_b200/variants/$v.so; fi
  echo "== $v"; timeout 600 python tools/stack_bench.py | tee gpurun_out/${TAG}_stack_bench_$v.json | cut -c1-700
  timeout 300 python -m pytest tests/test_gpu_stack.py -x -q 2>&1 | tail -1
done
