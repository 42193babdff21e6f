#!/bin/bash
# parity tests (all, no -x), micro-benchmarks, ncu capture of the contact-rich grasp step
TAG=${1:-q}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_pytest.log
python tools/quick_bench.py --tag $TAG 2>&1 | tail -1 | tee gpurun_out/${TAG}_quick.log
python tools/grasp_bench.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_grasp.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 130 -c 1 -o gpurun_out/${TAG}_grasp -f \
   python tools/profile_grasp.py 140 > gpurun_out/${TAG}_ncu_grasp.log 2>&1
