import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from gym_genesis_b200.model import franka_cube_stack_model
from gym_genesis_b200.sim import BatchedStackSim
from helpers import f32
g = np.load(os.path.join(ROOT, "tests", "golden", "stack_traj.npz"))
m = franka_cube_stack_model()
Q, QD, QA, T = g["q"], g["qd"], g["qacc"], g["act"]
sim = BatchedStackSim(m, Q.shape[0])
sim.set_state(f32(Q), f32(QD), f32(QA))
sim.step(torch.as_tensor(f32(T)))
nc, it = sim.step_info()
nc = nc.cpu().numpy()
bad = np.nonzero(nc != g["ncon"])[0]
print("mismatch idx", bad.tolist())
print("got", nc[bad].tolist())
print("exp", g["ncon"][bad].tolist())
WF = 327
work = torch.empty(sim.state.shape[1] * WF, dtype=torch.float32, device="cuda")
