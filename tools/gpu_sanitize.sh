#!/bin/bash
TAG=${1:-san}
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_stack.py > gpurun_out/${TAG}_memcheck.log 2>&1; tail -12 gpurun_out/${TAG}_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 20 python tools/sanitize_stack.py > gpurun_out/${TAG}_racecheck.log 2>&1; tail -15 gpurun_out/${TAG}_racecheck.log
