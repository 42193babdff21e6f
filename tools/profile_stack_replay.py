"""Replay the golden pick-and-stack rollout in every env up to a given step (for ncu captures of contact-rich stacking steps)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gym_genesis_b200.model import franka_cube_stack_model
from gym_genesis_b200.sim import BatchedStackSim
n = 16384
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 140
g = np.load(os.path.join(ROOT, "tests", "golden", "stack_traj.npz"))
m = franka_cube_stack_model()
sim = BatchedStackSim(m, n)
sim.reset(None, np.repeat(g["spawn"], n, axis=0))
act = torch.as_tensor(g["act_full"], device="cuda")
for k in range(steps):
    sim.step(act[k].repeat(n, 1).contiguous())
torch.cuda.synchronize()
print("contacts", sim.step_info()[0].float().mean().item())
