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
p.close()

# ---- round 2: coupled pick states (warp-per-env solver teams / the grid-wide coupled list) through BOTH step kernels, host I/O with
#      observation mirrors, IK with random re-starts
from gym_genesis_b200 import cabi  # noqa: E402
from helpers import pick_clutter_states  # noqa: E402

mp = franka_cube_pick_model()
gm = cabi.compile_gg(mp)
Qp, QDp, Tp = pick_clutter_states(mp, gm, np.random.default_rng(7), 150)
QAp = np.zeros_like(QDp)
for kern in ("roles", "mono"):
    os.environ["GG_STEP_KERNEL"] = kern
    ps = BatchedSim(mp, Qp.shape[0], seed=4, auto_reset=True, reset_prob=0.05)
    ps.set_state(f32(Qp), f32(QDp), f32(QAp))
    ap = torch.as_tensor(f32(Tp), device="cuda")
    for _ in range(3):
        ps.step(ap)
    hpin = torch.as_tensor(f32(Tp)).pin_memory()
    for _ in range(2):
        ps.step_host(hpin, host_obs=True)
    print(f"pick coupled ({kern}) ok", float(ps.get_state()[0].abs().sum()), float(ps.host_obs_agent.abs().sum()))
    ps.close()
os.environ.pop("GG_STEP_KERNEL", None)
pi = BatchedSim(mp, 96, seed=5)
pi.reset(None, None)
rng = np.random.RandomState(3)
tp = np.stack([rng.uniform(0.3, 0.9, 96), rng.uniform(-0.4, 0.4, 96), rng.uniform(0.05, 0.6, 96)], 1)      # some targets out of reach
tq = np.tile(np.array([0.0, 1.0, 0.0, 0.0]), (96, 1))
qik, eik = pi.inverse_kinematics(pos=f32(tp), quat=f32(tq), return_error=True, max_samples=50)
print("ik resample ok", float(qik.abs().sum()), float(eik.abs().sum()))
pi.close()
