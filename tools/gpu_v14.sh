#!/bin/bash
# v14 visit: final HEAD re-measure of the round (after the broad-phase refinements and the cube-kernel split) -- full GPU test tier, smoke, bench (both arms), launch list + one full capture of
# the headline step kernel.
TAG=${1:-v14}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -3 gpurun_out/${TAG}_smoke.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 300 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 600 python tools/stack_bench.py > gpurun_out/${TAG}_stack_bench_16k.json; cut -c1-400 gpurun_out/${TAG}_stack_bench_16k.json
timeout 600 python tools/stack_bench.py --envs 131072 --steps 60 --warmup 10 > gpurun_out/${TAG}_stack_bench_131k.json; cut -c1-400 gpurun_out/${TAG}_stack_bench_131k.json
python bench.py --impl reference --steps 50 --warmup 3 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err; tail -c 200 gpurun_out/${TAG}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv \
   python bench.py --steps 20 --warmup 10 --no-cpu-baseline --no-workload-b --no-workload-c --no-config4 > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 12 -c 1 -f -o gpurun_out/${TAG}_step python bench.py --steps 20 --warmup 10 --no-cpu-baseline --no-workload-b --no-workload-c --no-workload-d --no-config4 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stack_solve -s 12 -c 1 -f -o gpurun_out/${TAG}_stack_solve python tools/stack_bench.py --steps 20 --warmup 5 --skip-replay > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stack_prep -s 12 -c 1 -f -o gpurun_out/${TAG}_stack_prep python tools/stack_bench.py --steps 20 --warmup 5 --skip-replay > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stack_cubes -s 12 -c 1 -f -o gpurun_out/${TAG}_stack_cubes python tools/stack_bench.py --steps 20 --warmup 5 --skip-replay > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stack_finish -s 12 -c 1 -f -o gpurun_out/${TAG}_stack_finish python tools/stack_bench.py --steps 20 --warmup 5 --skip-replay > /dev/null 2>&1
ls gpurun_out | grep ${TAG}
