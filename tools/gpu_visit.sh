#!/bin/bash
# One iteration visit: parity tests, variant timings, full bench line.  Outputs -> gpurun_out/<tag>_*
TAG=${1:-visit}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_pytest.log
tools/gpu_variants.sh $TAG
python tools/grasp_bench.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_grasp.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 2500 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
