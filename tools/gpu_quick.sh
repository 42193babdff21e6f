#!/bin/bash
# Quick GPU visit while iterating on the kernel: parity tests + device-timed micro-benchmarks (workload A and the grasp replay).
TAG=${1:-q}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_pytest.log
python tools/quick_bench.py --tag $TAG 2>&1 | tail -2 | tee gpurun_out/${TAG}_quick.log
python tools/quick_bench.py --tag $TAG --envs 131072 --steps 50 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_quick.log
python tools/quick_bench.py --tag $TAG --envs 4096 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_quick.log
python tools/grasp_bench.py 2>&1 | tail -3 | tee gpurun_out/${TAG}_grasp.log
