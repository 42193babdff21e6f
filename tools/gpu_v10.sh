#!/bin/bash
# v10 visit: full GPU test tier, smoke, bench (both arms), CubeStack timings (default / 3-warp / 5-warp solver blocks),
# launch list of the bench command, full ncu captures of the CubeStack solve and prep kernels.  Outputs -> gpurun_out/<tag>_*
TAG=${1:-v10}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -6 gpurun_out/${TAG}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -3 gpurun_out/${TAG}_smoke.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 1800 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 50 --warmup 3 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err; tail -c 400 gpurun_out/${TAG}_bench_ref.json
for v in default stack_w3 stack_w5; do
  if [ $v = default ]; then unset GG_B200_LIB; else export GG_B200_LIB=$PWD/gym_genesis_b200/variants/$v.so; fi
  echo "== $v"; timeout 600 python tools/stack_bench.py | tee gpurun_out/${TAG}_stack_bench_$v.json | cut -c1-900
done
unset GG_B200_LIB
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv \
   python bench.py --steps 20 --warmup 10 --no-cpu-baseline --no-workload-b --no-workload-c > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stack_solve -s 12 -c 1 -f -o gpurun_out/${TAG}_stack_solve python tools/stack_bench.py --steps 20 --warmup 5 --skip-replay > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stack_prep -s 12 -c 1 -f -o gpurun_out/${TAG}_stack_prep python tools/stack_bench.py --steps 20 --warmup 5 --skip-replay > /dev/null 2>&1
ls gpurun_out | grep ${TAG}
