"""Developer micro-benchmark of the step kernel (same workload as bench.py, device-timed, flushed and L2-warm)."""
import argparse, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gym_genesis_b200.model import franka_cube_pick_model
from gym_genesis_b200.sim import BatchedSim

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=16384)
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--tag", default="")
a = ap.parse_args()
m = franka_cube_pick_model()
dev = torch.device("cuda", 0)
n = a.envs
sim = BatchedSim(m, n, device=dev, auto_reset=True, reset_prob=1 / 200., seed=0)
sim.reset(None, None)
g = torch.Generator(device=dev).manual_seed(1234)
home = torch.as_tensor(m.home_q, dtype=torch.float32, device=dev)
pool = home + 0.6 * (torch.rand((32, n, 9), generator=g, device=dev) - 0.5)
pool[:, :, 7:] = -0.02 + 0.06 * torch.rand((32, n, 2), generator=g, device=dev)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for k in range(50):
    sim.step(pool[k % 32])
torch.cuda.synchronize()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
for k in range(a.steps):
    flush.fill_(k & 0xFF)
    ev[k][0].record(); sim.step(pool[k % 32]); ev[k][1].record()
torch.cuda.synchronize()
cold = sum(x.elapsed_time(y) for x, y in ev) / a.steps
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for k in range(a.steps):
    sim.step(pool[k % 32])
e1.record(); torch.cuda.synchronize()
warm = e0.elapsed_time(e1) / a.steps
st = sim.stats().cpu().numpy()
print(json.dumps({"tag": a.tag, "envs": n, "cold_us": round(cold * 1e3, 2), "warm_us": round(warm * 1e3, 2),
                  "cold_Msteps": round(n / cold / 1e3, 1), "warm_Msteps": round(n / warm / 1e3, 1), "iters": round(st[6] / max(st[0], 1), 3)}))
