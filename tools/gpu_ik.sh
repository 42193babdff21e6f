#!/bin/bash
# GPU visit for the batched-IK row (8f-2): all GPU tests, IK timing, full bench line (incl. workload C), ncu capture of ik_kernel.
TAG=${1:-ik}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -25 gpurun_out/${TAG}_pytest.log
python tools/profile_ik.py 2>&1 | tail -1 | tee gpurun_out/${TAG}_ik_timing.log
python tools/profile_ik.py 131072 2>&1 | tail -1 | tee -a gpurun_out/${TAG}_ik_timing.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 1500 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ik_kernel -s 5 -c 1 -o gpurun_out/${TAG}_ik -f \
   python tools/profile_ik.py 16384 8 > gpurun_out/${TAG}_ncu_ik.log 2>&1
ls -la gpurun_out | tail -8
