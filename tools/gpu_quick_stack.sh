#!/bin/bash
# quick CubeStack iteration visit: stacking GPU tests + timing + launch list.  Outputs -> gpurun_out/<tag>_*
TAG=${1:-qs}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stack.py -x -q > gpurun_out/${TAG}_pytest_stack.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_stack.log
tail -4 gpurun_out/${TAG}_pytest_stack.log
timeout 600 python tools/stack_bench.py > gpurun_out/${TAG}_stack_bench.json 2> gpurun_out/${TAG}_stack_bench.err; cat gpurun_out/${TAG}_stack_bench.json; tail -3 gpurun_out/${TAG}_stack_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${TAG}_stack_launches.csv python tools/stack_bench.py --steps 20 --warmup 5 --skip-replay > /dev/null 2>&1
python - <<'PY'
import csv,collections,sys
lines=[l for l in open('gpurun_out/'+sys.argv[1]+'_stack_launches.csv') if not l.startswith('==')] if False else None
PY
grep -c stack_ gpurun_out/${TAG}_stack_launches.csv
