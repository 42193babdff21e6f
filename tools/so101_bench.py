"""SO-101 cube pick (BASELINE.json configs[3]) micro-benchmark: amplitude / friction sweep of the random-target workload, with the
episode statistics (Newton iterations, contacts) that explain the timing."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gym_genesis_b200.model_so101 import ISLAND_TOP_Z, so101_cube_pick_model  # noqa: E402
from gym_genesis_b200.sim import BatchedSim  # noqa: E402

dev = torch.device("cuda", 0)
n = 16384
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
CASES = ((0.3, 5.0, 1 / 200., True), (0.3, 5.0, 1 / 200., False), (0.3, 5.0, 0.0, True), (0.1, 5.0, 1 / 200., True), (0.3, 1.0, 1 / 200., True))
for amp, mu, reset_prob, mu_scaling in (CASES[:2] if "--one" in sys.argv else CASES):
    m = so101_cube_pick_model(success_height=ISLAND_TOP_Z + 0.1, pyramid_mu_scaling=mu_scaling)     # GENESIS_COMPAT_PYRAMID_MU_SCALING switch
    m.friction = mu
    sim = BatchedSim(m, n, device=dev, auto_reset=True, reset_prob=reset_prob, seed=0)
    sim.reset(None, None)
    g = torch.Generator(device=dev).manual_seed(4)
    home = torch.as_tensor(m.home_q.astype(np.float32), device=dev)
    pool = home + (torch.rand((8, n, 6), device=dev, generator=g) - 0.5) * 2 * amp
    for k in range(100):
        sim.step(pool[k % 8])
    sim.stats(clear=True)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(100)]
    for k in range(100):
        flush.fill_(k & 0xFF)
        ev[k][0].record(); sim.step(pool[k % 8]); ev[k][1].record()
    torch.cuda.synchronize()
    t = np.mean([a.elapsed_time(b) for a, b in ev])
    st = sim.stats().cpu().numpy()
    z = sim.get_state()[0][:, 8]
    print(json.dumps({"amp": amp, "mu": mu, "reset_prob": reset_prob, "pyramid_mu_scaling": mu_scaling,
                      "median_cube_depth_mm": round(float((m.plane_z + 0.02 - z.median()) * 1e3), 3), "us_per_step": round(t * 1e3, 1), "iters": round(st[6] / st[0], 3),
                      "contacts": round(st[7] / st[0], 3), "nonfinite": st[5]}))
    sim.close()
