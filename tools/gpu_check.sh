#!/bin/bash
# Light GPU visit: parity tests, smoke, bench (both arms). Outputs -> gpurun_out/<tag>_*
TAG=${1:-chk}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -12 gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 3000 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 50 --warmup 3 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err; tail -c 600 gpurun_out/${TAG}_bench_ref.json
