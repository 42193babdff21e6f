#!/bin/bash
TAG=${1:-bp}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/${TAG}_pytest.log
python tools/so101_bench.py 2>&1 | tail -4 | tee gpurun_out/${TAG}_so101_sweep.log
python tools/quick_bench.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_quick.log
python tools/grasp_bench.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_grasp.log
timeout 600 python tools/stack_bench.py | tee gpurun_out/${TAG}_stack_bench.json | cut -c1-600
