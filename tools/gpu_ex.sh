#!/bin/bash
# recorder / graph tests + the stacking collector example
TAG=${1:-ex}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_recorder.py tests/test_gpu_graph.py -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest.log
timeout 600 python examples/franka/stack_cube_state.py --num-envs 1024 --episodes 2 --out /tmp/stack_cube 2>&1 | tail -6 | tee gpurun_out/${TAG}_example_stack.log
timeout 600 python examples/franka/pick_cube_state.py --num-envs 1024 --episodes 1 --out /tmp/pick_cube 2>&1 | tail -3 | tee gpurun_out/${TAG}_example_pick.log
