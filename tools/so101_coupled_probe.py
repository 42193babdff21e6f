"""How many SO-101 envs take the coupled path per step?  Runs the random-target workload through the thread-per-env kernel, whose grid-wide
coupled list has a monotonic tail counter (gg_debug_ctl), and times both step kernels at two amplitudes.  Developer probe."""
import ctypes as C
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
for kern in ("mono", "roles"):
    os.environ["GG_STEP_KERNEL"] = kern
    for amp in (0.1, 0.2, 0.3):
        m = so101_cube_pick_model(success_height=ISLAND_TOP_Z + 0.1)
        sim = BatchedSim(m, n, device=dev, seed=0)
        sim.reset(None, None)
        g = torch.Generator(device=dev).manual_seed(4)
        home = torch.as_tensor(m.home_q.astype(np.float32), device=dev)
        pool = home + (torch.rand((8, n, 6), device=dev, generator=g) - 0.5) * 2 * amp
        for k in range(100):
            sim.step(pool[k % 8])
        torch.cuda.synchronize()
        ctl = (C.c_int * 4)()
        sim.lib.gg_debug_ctl.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        sim.lib.gg_debug_ctl(sim._h, ctl, 4)
        t0 = ctl[0]
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(50)]
        for k in range(50):
            flush.fill_(k & 0xFF)
            ev[k][0].record(); sim.step(pool[k % 8]); ev[k][1].record()
        torch.cuda.synchronize()
        sim.lib.gg_debug_ctl(sim._h, ctl, 4)
        t = np.mean([a.elapsed_time(b) for a, b in ev])
        q = sim.get_state()[0]
        lo = torch.as_tensor(np.array(list(sim.gg_model.lim_lo)[:6], np.float32), device=dev)
        hi = torch.as_tensor(np.array(list(sim.gg_model.lim_hi)[:6], np.float32), device=dev)
        near = ((q[:, :6] < lo + 0.05) | (q[:, :6] > hi - 0.05)).float().mean(0).cpu().numpy().round(3).tolist()
        st = sim.stats().cpu().numpy()
        print(json.dumps({"kernel": kern, "amp": amp, "us_per_step": round(t * 1e3, 1), "coupled_per_step": (ctl[0] - t0) / 50.0,
                          "iters": round(st[6] / max(st[0], 1), 3), "joints_near_limit": near}))
        sim.close()
