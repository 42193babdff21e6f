#!/bin/bash
# v16 visit: HEAD of round 1 -- sanitizer, full GPU tests, smoke, bench (both arms), stack timings, launch list.
TAG=${1:-v16}
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_stack.py > gpurun_out/${TAG}_memcheck.log 2>&1; tail -3 gpurun_out/${TAG}_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 20 python tools/sanitize_stack.py > gpurun_out/${TAG}_racecheck.log 2>&1; tail -2 gpurun_out/${TAG}_racecheck.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -3 gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 200 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 50 --warmup 3 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err; tail -c 200 gpurun_out/${TAG}_bench_ref.json
timeout 600 python tools/stack_bench.py > gpurun_out/${TAG}_stack_bench_16k.json; cut -c1-300 gpurun_out/${TAG}_stack_bench_16k.json
timeout 600 python tools/stack_bench.py --envs 131072 --steps 60 --warmup 10 > gpurun_out/${TAG}_stack_bench_131k.json; cut -c1-300 gpurun_out/${TAG}_stack_bench_131k.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv \
   python bench.py --steps 20 --warmup 10 --no-cpu-baseline --no-workload-b --no-workload-c --no-config4 > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stack_solve -s 12 -c 1 -f -o gpurun_out/${TAG}_stack_solve python tools/stack_bench.py --steps 20 --warmup 5 --skip-replay > /dev/null 2>&1
ls gpurun_out | grep ${TAG}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stack_solve -s 130 -c 1 -f -o gpurun_out/${TAG}_stack_solve_grasp python tools/profile_stack_replay.py 140 > /dev/null 2>&1
timeout 300 python examples/franka/stack_cube_state.py --num-envs 1024 --episodes 1 --scene kitchen --out /tmp/stack_k 2>&1 | tail -2 | tee gpurun_out/${TAG}_example_stack_kitchen.log
ls gpurun_out | grep ${TAG} | wc -l
