"""Timing of the CubeStack step path on one GPU: CUDA events around gg_stack_step, L2 flushed between steps.
  workload S1: random joint targets resampled every step, Bernoulli(1/200) resets + auto-reset  (BASELINE configs[2] shape)
  workload S2: open-loop replay of the golden pick-and-stack rollout in every env (contact rich: 16-52 contacts)
Prints one JSON line."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gym_genesis_b200.model import FRANKA_HOME, franka_cube_stack_model  # noqa: E402
from gym_genesis_b200.sim import BatchedStackSim  # noqa: E402


def timed(fn, steps, flush):
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for k in range(steps):
        flush.fill_(k & 0xFF)
        ev[k][0].record()
        fn(k)
        ev[k][1].record()
    torch.cuda.synchronize()
    return np.array([a.elapsed_time(b) for a, b in ev])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--skip-replay", action="store_true")
    a = ap.parse_args()
    n = a.envs
    m = franka_cube_stack_model()
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {"envs": n}
    sim = BatchedStackSim(m, n, auto_reset=True, reset_prob=1.0 / 200, seed=0)
    sim.reset(None, None)
    g = torch.Generator(device=dev).manual_seed(0)
    home = torch.as_tensor(FRANKA_HOME.astype(np.float32), device=dev)
    acts = home + (torch.rand((a.steps + a.warmup, n, 9), device=dev, generator=g) - 0.5) * 0.6
    acts[:, :, 7:] = torch.rand((a.steps + a.warmup, n, 2), device=dev, generator=g) * 0.06 - 0.02
    for k in range(a.warmup):
        sim.step(acts[k])
    sim.stats(clear=True)
    ms = timed(lambda k: sim.step(acts[a.warmup + k]), a.steps, flush)
    st = sim.stats().cpu().numpy()
    out["s1"] = {"ms_per_step": float(ms.mean()), "env_steps_per_s": n / (ms.mean() * 1e-3), "mean_contacts": st[7] / max(st[0], 1),
                 "mean_iters": st[6] / max(st[0], 1), "nonfinite": st[5], "p50_ms": float(np.median(ms)), "max_ms": float(ms.max())}
    sim.close()
    if not a.skip_replay:
        g = np.load(os.path.join(ROOT, "tests", "golden", "stack_traj.npz"))
        act = torch.as_tensor(g["act_full"], device=dev)
        sim = BatchedStackSim(m, n)
        sim.reset(None, np.repeat(g["spawn"], n, axis=0))
        rows = [act[k].repeat(n, 1).contiguous() for k in range(act.shape[0])]
        sim.stats(clear=True)
        ms = timed(lambda k: sim.step(rows[k]), act.shape[0], flush)
        st = sim.stats().cpu().numpy()
        out["s2"] = {"ms_per_step": float(ms.mean()), "env_steps_per_s": n / (ms.mean() * 1e-3), "mean_contacts": st[7] / max(st[0], 1),
                     "mean_iters": st[6] / max(st[0], 1), "success_rate": float(sim.reward.mean().item()),
                     "ms_by_stage": [float(ms[i:j].mean()) for i, j in ((0, 40), (40, 80), (80, 110), (110, 160), (160, 230), (230, 290), (290, 320), (320, 350))]}
        sim.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
