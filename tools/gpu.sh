#!/bin/bash
# ONE parametrised GPU visit script (replaces the per-visit gpu_v*.sh / gpu_q*.sh copies of round 1).
#   tools/gpu.sh TAG stage [stage ...]          outputs -> gpurun_out/<TAG>_*
# stages:
#   build        recompile every CUDA unit on the box (__graft_entry__.build(force=True)): proves the compile where it runs
#   pytest       full GPU test tier                         pytest:<expr>  only tests matching -k <expr>
#   newtests     the full-size parity tests only
#   smoke        __graft_entry__.smoke()
#   quick        tools/quick_bench.py at 16384 / 131072 / 4096 envs + grasp replay
#   variants     every gym_genesis_b200/variants/*.so through quick_bench at the env counts in $ENVS (+ grasp replay when GRASP=1)
#   bench        bench.py (own arm) + reference arm
#   benchlite    bench.py --no-workload-b --no-workload-c --no-config4 --no-workload-d (headline legs only)
#   launches     ncu launch list of bench.py
#   ncu_mono     ncu --set full of the thread-per-env step kernel at 131072 envs (BASELINE configs[4] on one GPU)
#   ncu_step     ncu --set full of the step kernel (bench workload; launch 212 = steady state after 200 warm-up steps: the fingers ride their
#                stops only after a few dozen steps, earlier launches skip the limit-row solve)      ncu_grasp: contact-rich step 130 of the grasp replay
#   ncu_stack    ncu --set full of the stacking solver (random workload)
#   stack        tools/stack_bench.py at 16384 / 131072 envs
#   sanitize     compute-sanitizer memcheck + racecheck on the pick and stack kernels
#   multi / c5   NG=N: the driver's torchrun launch line on N GPUs (run under gpurun --gpus N): default workload (weak scaling) /
#                BASELINE configs[4] --total-envs 131072 (strong scaling); headline legs only
TAG=${1:-run}; shift
mkdir -p gpurun_out
O=gpurun_out/${TAG}
for stage in "$@"; do
  arg=${stage#*:}; name=${stage%%:*}
  echo "=== stage ${stage}"
  case $name in
    build)    python -c "import __graft_entry__ as g; g.build(force=True); print('build(force) ok')" > ${O}_build.log 2>&1; tail -2 ${O}_build.log ;;
    pytest)   if [ "$arg" != "$name" ]; then K=(-k "$arg"); else K=(); fi
              timeout 2400 python -m pytest tests -m gpu -x -q "${K[@]}" > ${O}_pytest.log 2>&1; echo "pytest exit $?" >> ${O}_pytest.log; tail -6 ${O}_pytest.log ;;
    newtests) timeout 1800 python -m pytest tests/test_gpu_fullsize_parity.py -m gpu -x -q > ${O}_newtests.log 2>&1; echo "pytest exit $?" >> ${O}_newtests.log; tail -25 ${O}_newtests.log ;;
    smoke)    python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; tail -2 ${O}_smoke.log ;;
    quick)    python tools/quick_bench.py --tag $TAG 2>&1 | tail -1 | tee ${O}_quick.log
              python tools/quick_bench.py --tag $TAG --envs 131072 --steps 50 2>&1 | tail -1 | tee -a ${O}_quick.log
              python tools/quick_bench.py --tag $TAG --envs 4096 2>&1 | tail -1 | tee -a ${O}_quick.log
              python tools/grasp_bench.py 2>&1 | tail -3 | tee ${O}_grasp.log ;;
    variants) for so in gym_genesis_b200/variants/*.so; do
                v=$(basename $so .so)
                for n in ${ENVS:-16384}; do GG_B200_LIB=$PWD/$so python tools/quick_bench.py --tag $v --envs $n --steps ${STEPS:-200} 2>&1 | tail -1 | tee -a ${O}_variants.log; done
                if [ -n "$GRASP" ]; then GG_B200_LIB=$PWD/$so python tools/grasp_bench.py 2>&1 | tail -1 | tee -a ${O}_variants.log; fi
              done ;;
    bench)    python bench.py > ${O}_bench.json 2> ${O}_bench.err; tail -c 1500 ${O}_bench.json; tail -3 ${O}_bench.err
              python bench.py --impl reference --steps 20 --warmup 3 > ${O}_bench_ref.json 2>> ${O}_bench.err; tail -c 400 ${O}_bench_ref.json ;;
    benchlite) python bench.py --no-workload-b --no-workload-c --no-config4 --no-workload-d > ${O}_benchlite.json 2> ${O}_benchlite.err; tail -c 1500 ${O}_benchlite.json; tail -3 ${O}_benchlite.err ;;
    launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file ${O}_launches.csv \
                python bench.py --steps 20 --warmup 10 --no-cpu-baseline --no-workload-b --no-workload-c --no-config4 --no-config2 > ${O}_ncu_bench.log 2>&1; wc -l ${O}_launches.csv ;;
    ncu_step) timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 212 -c 1 -f -o ${O}_step \
                python bench.py --steps 20 --warmup 200 --no-cpu-baseline --no-workload-b --no-workload-c --no-config4 --no-config2 --no-workload-d > ${O}_ncu_full.log 2>&1; ls -la ${O}_step.ncu-rep ;;
    ncu_mono) timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_mono_kernel -s 212 -c 1 -f -o ${O}_step_mono \
                python bench.py --total-envs 131072 --steps 20 --warmup 200 --no-cpu-baseline --no-workload-b --no-workload-c --no-config4 --no-config2 --no-workload-d > ${O}_ncu_mono.log 2>&1; ls -la ${O}_step_mono.ncu-rep ;;
    ncu_grasp) timeout 600 ncu --set full --clock-control none --import-source on -k regex:"step_kernel|coupled" -s 130 -c 1 -f -o ${O}_grasp \
                python tools/profile_grasp.py 140 > ${O}_ncu_grasp.log 2>&1; ls -la ${O}_grasp.ncu-rep ;;
    ncu_stack) timeout 900 ncu --set full --clock-control none --import-source on -k regex:stack_solve -s 12 -c 1 -f -o ${O}_stack_solve \
                python tools/stack_bench.py --steps 20 --warmup 5 --skip-replay > /dev/null 2>&1; ls -la ${O}_stack_solve.ncu-rep ;;
    stack)    timeout 600 python tools/stack_bench.py > ${O}_stack_bench_16k.json; cut -c1-300 ${O}_stack_bench_16k.json
              timeout 600 python tools/stack_bench.py --envs 131072 --steps 60 --warmup 10 > ${O}_stack_bench_131k.json; cut -c1-300 ${O}_stack_bench_131k.json ;;
    sanitize) timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_stack.py > ${O}_memcheck.log 2>&1; tail -3 ${O}_memcheck.log
              timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 20 python tools/sanitize_stack.py > ${O}_racecheck.log 2>&1; tail -2 ${O}_racecheck.log ;;
    multi)    NG=${NG:-2}; LITE="--no-workload-b --no-workload-c --no-config4 --no-workload-d --no-config2 --no-cpu-baseline"
              timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $NG $LITE \
                > ${O}_bench_${NG}gpu.json 2> ${O}_bench_${NG}gpu.err; tail -c 900 ${O}_bench_${NG}gpu.json; tail -2 ${O}_bench_${NG}gpu.err ;;
    c5)       NG=${NG:-2}; LITE="--no-workload-b --no-workload-c --no-config4 --no-workload-d --no-config2 --no-cpu-baseline"
              timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $NG --total-envs 131072 $LITE \
                > ${O}_c5_${NG}gpu.json 2> ${O}_c5_${NG}gpu.err; tail -c 900 ${O}_c5_${NG}gpu.json; tail -2 ${O}_c5_${NG}gpu.err ;;
    *)        echo "unknown stage $stage" ;;
  esac
done
ls gpurun_out | grep "^${TAG}_" | wc -l
