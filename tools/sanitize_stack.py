"""Small CubeStack / CubePick runs for compute-sanitizer (memcheck / racecheck): golden-trajectory states (coupled solves, 16-52
contacts), clustered states (all 39 dofs active), random resets, a few steps each."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gym_genesis_b200.model import FRANKA_HOME, franka_cube_pick_model, franka_cube_stack_model  # noqa: E402
from gym_genesis_b200.model_so101 import so101_cube_stack_model  # noqa: E402
from gym_genesis_b200.sim import BatchedSim, BatchedStackSim  # noqa: E402
from helpers import f32, stack_cluster_states  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "stack_traj.npz"))
m = franka_cube_stack_model()
idx = np.arange(0, len(g["q"]), 3)
Qc, QDc, QAc, Tc = stack_cluster_states(g)
Q = np.concatenate([g["q"][idx], Qc]); QD = np.concatenate([g["qd"][idx], QDc]); QA = np.concatenate([g["qacc"][idx], QAc])
T = np.concatenate([g["act"][idx], Tc])
sim = BatchedStackSim(m, Q.shape[0], auto_reset=True, reset_prob=0.05, seed=1)
sim.set_state(f32(Q), f32(QD), f32(QA))
a = torch.as_tensor(f32(T), device="cuda")
for _ in range(3):
    sim.step(a)
sim.reset(None, None)
sim.step(a)
print("franka stack ok", float(sim.get_state()[0].abs().sum()))
sim.close()
ms = so101_cube_stack_model()
s2 = BatchedStackSim(ms, 70, seed=2)
s2.reset(None, None)
h = torch.as_tensor(f32(np.tile(ms.home_q, (70, 1))), device="cuda")
for _ in range(3):
    s2.step(h)
print("so101 stack ok", float(s2.get_state()[0].abs().sum()))
s2.close()
p = BatchedSim(franka_cube_pick_model(), 100, seed=3)
p.reset(None, None)
hp = torch.as_tensor(f32(np.tile(FRANKA_HOME, (100, 1))), device="cuda")
for _ in range(3):
    p.step(hp)
print("pick ok", float(p.get_state()[0].abs().sum()))
