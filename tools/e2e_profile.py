"""Where does the host time of one e2e step go?  cProfile over the README loop (pinned host action in, host results out) for the
default (fresh result tensors) and alias_outputs=True modes.  Developer tool; prints the top functions by own time."""
import cProfile
import io
import os
import pstats
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gym_genesis_b200  # noqa: F401,E402
from gym_genesis_b200.gym_api import gym  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
steps = 3000
dev = "cuda:0"
rng = np.random.RandomState(0)
from gym_genesis_b200.model import franka_cube_pick_model  # noqa: E402
home = franka_cube_pick_model().home_q.astype(np.float32)
# bench.py's action distribution: random joint targets within +-0.3 rad of the home pose
acts = [torch.as_tensor((home[None] + rng.uniform(-0.3, 0.3, (n, 9))).astype(np.float32)).pin_memory() for _ in range(4)]


def loop(env, k0, k1):
    chk = 0.0
    for k in range(k0, k1):
        obs, reward, terminated, truncated, info = env.step(acts[k % 4])
        r_, t_ = reward.numpy(), terminated.numpy()
        chk += r_[0] + t_[-1] + r_[n // 2]
    return chk


def loop_full(env, k0, k1):
    chk = 0.0
    for k in range(k0, k1):
        obs, reward, terminated, truncated, info = env.step(acts[k % 4])
        r_, t_ = reward.numpy(), terminated.numpy()
        chk += r_[0] + t_[-1] + r_[n // 2] + obs["agent_pos"].numpy()[n // 2, 0] + obs["environment_state"].numpy()[-1, 10]
    return chk


if "--full" in sys.argv:
    for rep in range(3):
        env = gym.make("gym_genesis/CubePick-v0", robot="franka", num_envs=n, device=dev, auto_reset=True, reset_prob=0.005, seed=1,
                       host_io=True, host_obs=True).unwrapped
        env.reset(seed=1)
        loop_full(env, 0, 200)
        t0 = time.perf_counter()
        loop_full(env, 0, steps)
        t1 = time.perf_counter()
        print(f"full-observation mode: {(t1 - t0) / steps * 1e6:.1f} us / step  ({n * steps / (t1 - t0) / 1e6:.1f} M env-steps/s)")
        env.close()
    sys.exit(0)

for kw in ({}, {"alias_outputs": True}):
    env = gym.make("gym_genesis/CubePick-v0", robot="franka", num_envs=n, device=dev, auto_reset=True, reset_prob=0.005, seed=1,
                   host_io=True, **kw).unwrapped
    env.reset(seed=1)
    loop(env, 0, 200)
    t0 = time.perf_counter()
    loop(env, 0, steps)
    t1 = time.perf_counter()
    print(f"mode {kw}: {(t1 - t0) / steps * 1e6:.1f} us / step  ({n * steps / (t1 - t0) / 1e6:.1f} M env-steps/s)")
    pr = cProfile.Profile()
    pr.enable()
    loop(env, 0, steps)
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(14)
    print("\n".join(l for l in s.getvalue().splitlines() if l.strip())[:3500])
    env.close()

# ---- breakdown on the raw simulator: launch cost (wait=False), kernel time with host I/O (events), blocking step
from gym_genesis_b200.sim import BatchedSim  # noqa: E402

sim = BatchedSim(franka_cube_pick_model(), n, device=dev, seed=1, auto_reset=True, reset_prob=0.005)
sim.reset()
dacts = [a.cuda() for a in acts]
for k in range(200):
    sim.step_host(acts[k % 4])
torch.cuda.synchronize()
t0 = time.perf_counter()
for k in range(steps):
    sim.step_host(acts[k % 4])
t1 = time.perf_counter()
print(f"step_host wait=True            : {(t1 - t0) / steps * 1e6:.1f} us / step")
torch.cuda.synchronize()
t0 = time.perf_counter()
for k in range(steps):
    sim.step_host(acts[k % 4], wait=False)
t_launch = time.perf_counter()
torch.cuda.synchronize()
t1 = time.perf_counter()
print(f"step_host wait=False           : {(t_launch - t0) / steps * 1e6:.1f} us / call issued, {(t1 - t0) / steps * 1e6:.1f} us / step drained")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for k in range(steps):
    sim.step_host(acts[k % 4], wait=False)
ev[1].record()
torch.cuda.synchronize()
print(f"step_host wait=False (events)  : {ev[0].elapsed_time(ev[1]) / steps * 1e3:.1f} us / step on the device")
ev[0].record()
for k in range(steps):
    sim.step(dacts[k % 4])
ev[1].record()
torch.cuda.synchronize()
print(f"step (device action, events)   : {ev[0].elapsed_time(ev[1]) / steps * 1e3:.1f} us / step on the device")
t0 = time.perf_counter()
for k in range(steps):
    sim.step(dacts[k % 4])
    torch.cuda.current_stream().synchronize()
t1 = time.perf_counter()
print(f"step + stream.synchronize      : {(t1 - t0) / steps * 1e6:.1f} us / step")
import subprocess  # noqa: E402
print(subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,pstate", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip())
print("cpu count", os.cpu_count(), open("/proc/cpuinfo").read().split("model name")[1].split("\n")[0])
