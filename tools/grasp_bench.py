"""Device-timed replay of the scripted-grasp fixture on 16384 envs (workload B of bench.py): ms per step by stage."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gym_genesis_b200.model import franka_cube_pick_model
from gym_genesis_b200.sim import BatchedSim
d = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "grasp_traj.npz"))
K = d["xy"].shape[0]; n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384; rep = n // K
sim = BatchedSim(franka_cube_pick_model(), K * rep)
acts = torch.as_tensor(np.tile(d["actions"], (rep, 1, 1)), device="cuda").transpose(0, 1).contiguous()
xy = np.tile(d["xy"], (rep, 1))
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for w in range(2):
    sim.reset(None, xy)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(200)]
    for t in range(200):
        flush.fill_(t & 0xFF)
        ev[t][0].record(); sim.step(acts[t]); ev[t][1].record()
    torch.cuda.synchronize()
tb = np.array([a.elapsed_time(b) for a, b in ev])
st = sim.stats().cpu().numpy()
print(json.dumps({"envs": K * rep, "Msteps": round(K * rep * 200 / tb.sum() / 1e3, 2), "ms_by_stage": [round(float(tb[i * 40:(i + 1) * 40].mean()), 4) for i in range(5)],
                  "iters": round(st[6] / st[0], 3), "contacts": round(st[7] / st[0], 2), "success": float(sim.terminated.float().mean().item()), "nan": st[5]}))
