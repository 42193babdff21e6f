"""ncu / timing target for the batched IK kernel: 16384 envs after reset, hover targets (cube + 0.115 m, hand down)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_genesis_b200.model import franka_cube_pick_model  # noqa: E402
from gym_genesis_b200.sim import BatchedSim  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
sim = BatchedSim(franka_cube_pick_model(), n)
sim.reset(None, None)
dev = sim.device
pos = sim.obs_env[:, :3] + torch.tensor([0.0, 0.0, 0.115], device=dev)
quat = torch.tensor([0.0, 1.0, 0.0, 0.0], device=dev).expand(n, -1).contiguous()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
for k in range(reps):
    flush.fill_(k)
    ev[k][0].record()
    q, err, it = sim.inverse_kinematics(pos=pos, quat=quat, return_error=True, return_iters=True)
    ev[k][1].record()
torch.cuda.synchronize()
t = sorted(a.elapsed_time(b) for a, b in ev[2:])
print(f"ik_kernel n={n}: median {t[len(t) // 2] * 1e3:.1f} us / launch (incl. 3 torch.empty), {n / (t[len(t) // 2] * 1e-3) / 1e6:.1f} M solves/s, "
      f"mean updates {it.float().mean().item():.2f}, max {it.max().item()}, solved {(err[:, 0] < 5e-4).float().mean().item():.3f}")
