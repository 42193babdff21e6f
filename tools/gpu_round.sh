#!/bin/bash
# One GPU visit: parity tests, smoke, bench, launch list, full ncu capture of the step kernel. Outputs -> gpurun_out/<tag>_*
TAG=${1:-run}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 3000 gpurun_out/${TAG}_bench.json
python bench.py --impl reference --steps 50 --warmup 3 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err; tail -c 600 gpurun_out/${TAG}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv \
   python bench.py --steps 20 --warmup 10 --no-cpu-baseline --no-workload-b --no-workload-c > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 12 -c 1 -o gpurun_out/${TAG}_step -f \
   python bench.py --steps 20 --warmup 10 --no-cpu-baseline --no-workload-b --no-workload-c > gpurun_out/${TAG}_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 130 -c 1 -o gpurun_out/${TAG}_grasp -f \
   python tools/profile_grasp.py 140 > gpurun_out/${TAG}_ncu_grasp.log 2>&1
ls -la gpurun_out
