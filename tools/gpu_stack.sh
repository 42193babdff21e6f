#!/bin/bash
# GPU visit for the CubeStack path: parity tests, timing, launch list + one full ncu capture of the solve kernel.
TAG=${1:-stk}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stack.py -x -q > gpurun_out/${TAG}_pytest_stack.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_stack.log
tail -15 gpurun_out/${TAG}_pytest_stack.log
timeout 600 python tools/stack_bench.py > gpurun_out/${TAG}_stack_bench.json 2> gpurun_out/${TAG}_stack_bench.err; cat gpurun_out/${TAG}_stack_bench.json; tail -3 gpurun_out/${TAG}_stack_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${TAG}_stack_launches.csv python tools/stack_bench.py --steps 20 --warmup 5 --skip-replay > /dev/null 2>&1
grep -c stack_ gpurun_out/${TAG}_stack_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stack_solve -s 12 -c 1 -o gpurun_out/${TAG}_stack_solve python tools/stack_bench.py --steps 20 --warmup 5 --skip-replay > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stack_prep -s 12 -c 1 -o gpurun_out/${TAG}_stack_prep python tools/stack_bench.py --steps 20 --warmup 5 --skip-replay > /dev/null 2>&1
ls -la gpurun_out/ | tail -8
