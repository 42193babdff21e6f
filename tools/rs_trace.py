"""Timeline of the role-per-SM kernel (library built with -DGG_RS_TRACE): per group, nanoseconds from the earliest milestone of the launch."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gym_genesis_b200.model import franka_cube_pick_model
from gym_genesis_b200.sim import BatchedSim
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
m = franka_cube_pick_model()
sim = BatchedSim(m, n, seed=0, auto_reset=True, reset_prob=1 / 200.)
sim.reset(None, None)
g = torch.Generator(device="cuda").manual_seed(1)
home = torch.as_tensor(m.home_q, dtype=torch.float32, device="cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for k in range(30):
    act = home + 0.6 * (torch.rand((n, 9), generator=g, device="cuda") - 0.5)
    flush.fill_(k)
    sim.step(act)
torch.cuda.synchronize()
buf = (C.c_int * 8192)()
sim.lib.gg_debug_ctl.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
assert sim.lib.gg_debug_ctl(sim._h, buf, 8192) == 0
T = np.frombuffer(buf, dtype=np.uint64, offset=64 * 4)[: 64 * 3 * 16].reshape(64, 3, 16).astype(np.int64)
ng = min(64, (n + 31) // 32)
t0 = T[:ng, :, 0].min()
names = {0: ["start", "crba+chol+lim", "FS arrived", "solved+ROB posted", "eef", "C1 arrived", "end", "|crba", "chol", "chol_solve", "block_solve"],
         1: ["start", "rnea", "C0 arrived", "FS posted"],
         2: ["start", "C0 posted", "solved", "FS arrived", "integrated", "ROB arrived", "end"]}
for gi in (0, 1, 2, 49, 50, 63):
    if gi >= ng: continue
    for r, rn in ((0, "MAIN"), (1, "RNEA"), (2, "CUBE")):
        k = len(names[r])
        ts = (T[gi, r, :k] - t0) / 1000.0
        print(f"group {gi:3d} {rn} sm {int(T[gi, r, 15]):3d}: " + "  ".join(f"{nm} {t:6.2f}" for nm, t in zip(names[r], ts)))
end = (T[:ng, 0, 6].max() - t0) / 1000.0
print(f"first start -> last MAIN end over the first {ng} groups: {end:.2f} us; starts spread {((T[:ng,:,0].max()-t0)/1000.0):.2f} us")
