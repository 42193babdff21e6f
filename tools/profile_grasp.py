"""Replays the scripted-grasp fixture on 16384 envs (workload B) -- used under ncu to profile a contact-rich step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gym_genesis_b200.model import franka_cube_pick_model
from gym_genesis_b200.sim import BatchedSim
d = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "grasp_traj.npz"))
K = d["xy"].shape[0]; rep = 16384 // K
sim = BatchedSim(franka_cube_pick_model(), K * rep)
acts = torch.as_tensor(np.tile(d["actions"], (rep, 1, 1)), device="cuda").transpose(0, 1).contiguous()
sim.reset(None, np.tile(d["xy"], (rep, 1)))
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
for t in range(steps):
    sim.step(acts[t])
torch.cuda.synchronize()
st = sim.stats().cpu().numpy()
print("iters/env-step", st[6] / st[0], "contacts/env-step", st[7] / st[0])
