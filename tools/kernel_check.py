"""Developer check: the thread-per-env (mono) step kernel against the role kernel on the same inputs (state bit-identical expected for the
register path; observations differ by the rounding of the unsplit end-effector chain)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gym_genesis_b200.model import franka_cube_pick_model
from gym_genesis_b200.sim import BatchedSim

m = franka_cube_pick_model()
for n in (64, 1000, 16384):
    sims = []
    for kern in ("roles", "mono"):
        os.environ["GG_STEP_KERNEL"] = kern
        s = BatchedSim(m, n, seed=5, auto_reset=True, reset_prob=0.02, time_limit=True)
        s.reset(None, None)
        sims.append(s)
    torch.cuda.synchronize()
    print(f"n={n}: reset done", flush=True)
    g = torch.Generator(device="cuda").manual_seed(1)
    home = torch.as_tensor(m.home_q, dtype=torch.float32, device="cuda")
    for k in range(20):
        act = home + 0.6 * (torch.rand((n, 9), generator=g, device="cuda") - 0.5)
        act[:, 7:] = -0.02 + 0.06 * torch.rand((n, 2), generator=g, device="cuda")
        for s in sims:
            s.step(act)
        torch.cuda.synchronize()
    a, b = sims
    qa, qb = a.get_state(), b.get_state()
    dq = max((x - y).abs().max().item() for x, y in zip(qa, qb))
    print(f"n={n}: |state diff| {dq:.3e}  obs_agent {((a.obs_agent - b.obs_agent).abs().max().item()):.3e} obs_env {((a.obs_env - b.obs_env).abs().max().item()):.3e} "
          f"flags equal {bool(torch.equal(a.terminated, b.terminated) and torch.equal(a.truncated, b.truncated))} counters equal "
          f"{bool(torch.equal(a.step_count, b.step_count) and torch.equal(a.episode, b.episode))} stats {a.stats().cpu().numpy().tolist()} {b.stats().cpu().numpy().tolist()}", flush=True)
    for s in sims:
        s.close()
print("rs_check done")
