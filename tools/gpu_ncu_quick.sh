#!/bin/bash
# ncu --set full capture of one step-kernel launch of tools/quick_bench.py at a given env count: tools/gpu_ncu_quick.sh TAG ENVS [kernel regex]
TAG=$1; ENVS=${2:-16384}; K=${3:-step_kernel}
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 60 -c 1 -f -o gpurun_out/${TAG}_step_${ENVS} \
   python tools/quick_bench.py --steps 20 --envs $ENVS > gpurun_out/${TAG}_ncu_${ENVS}.log 2>&1
ls -la gpurun_out/${TAG}_step_${ENVS}.ncu-rep
