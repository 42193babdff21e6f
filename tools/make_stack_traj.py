"""Generates tests/golden/stack_traj.npz: a scripted pick-and-stack rollout of the CubeStack scene in the fp64 ORACLE
(test infrastructure), sampled every 3rd step, with the oracle's own next state stored beside every sample.

The expert restates the stage machine of the reference collectors (examples/franka/stack_cube_state.py: hover, descend,
grasp, lift, carry over cube_2, lower, release, retreat) with a numpy damped-least-squares IK standing in for
robot.inverse_kinematics.  Run from the repo root:  python tools/make_stack_traj.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from gym_genesis_b200.model import FRANKA_HOME, franka_cube_stack_model, quat_to_mat  # noqa: E402
from helpers import f32, ik_numpy  # noqa: E402
from oracle.binding import Oracle  # noqa: E402

# (reference frame, height of the hand above it, gripper target, steps)
PLAN = [("c1", 0.2, 0.04, 40), ("c1", 0.105, 0.04, 40), ("c1", 0.105, -0.02, 30), ("lift", 0.25, -0.02, 50),
        ("c2", 0.23, -0.02, 70), ("c2", 0.148, -0.02, 60), ("c2", 0.148, 0.04, 30), ("c2", 0.3, 0.04, 30)]
SPAWN = np.array([[[0.55, 0.05], [0.5, -0.15], [0.35, 0.25], [0.65, 0.25], [0.4, -0.25]]])


def main(stride=3):
    m = franka_cube_stack_model()
    o = Oracle(m, np.float64)
    o32 = Oracle(m, np.float32)
    Rdown = quat_to_mat([0, 1, 0, 0])
    Q, QD, QA = o.new_state(1)
    o.reset(Q, QD, QA, SPAWN)
    o.step(Q, QD, QA, FRANKA_HOME.copy()[None])
    rec = {k: [] for k in ("q", "qd", "qacc", "act", "qn", "qdn", "ncon", "asm")}
    k, hold, spread = 0, None, 0.0
    act_full = []
    for name, dz, grip, steps in PLAN:
        for _ in range(steps):
            c1, c2 = Q[0, 9:12].copy(), Q[0, 16:19].copy()
            if name == "c1":
                tp, hold = c1 + [0, 0, dz], c1.copy()
            elif name == "lift":
                tp = hold + [0, 0, dz]
            else:
                tp = np.array([c2[0], c2[1], 0.02 + dz])
            qt = ik_numpy(m, tp, Rdown, Q[0, :9].copy())
            act = np.concatenate([qt[:7], [grip, grip]])[None]
            if k % stride == 0:
                # samples are rounded to float32 so that the fp32 CUDA path and the fp64 oracle start from identical bits
                q, qd, qa, a = (f32(x).astype(np.float64) for x in (Q, QD, QA, act))
                qn, qdn, qan = q.copy(), qd.copy(), qa.copy()
                info, _ = o.step(qn, qdn, qan, a, iterations=100, tol=1e-12)
                q3, qd3, qa3 = f32(q).copy(), f32(qd).copy(), f32(qa).copy()
                o32.step(q3, qd3, qa3, f32(a), iterations=100, tol=1e-7)
                spread = max(spread, np.abs(qd3 - qdn).max())
                asm = np.abs(o.dynamics(q[0], qd[0], a[0])["a_smooth"]).max() * 0.01
                for key, v in zip(rec, (q[0], qd[0], qa[0], a[0], qn[0], qdn[0], info[0, 0], asm)):
                    rec[key].append(v)
            act_full.append(f32(act[0]))
            o.step(Q, QD, QA, act, iterations=100, tol=1e-10)
            k += 1
    agent, envs, reward, term = o.obs(Q)
    out = {key: np.asarray(v) for key, v in rec.items()}
    out["act_full"] = np.asarray(act_full)          # every step's action (float32): open-loop replay in tools/stack_bench.py
    out["spawn"] = SPAWN.astype(np.float32)
    out["final_reward"] = reward
    out["final_q"] = Q
    path = os.path.join(ROOT, "tests", "golden", "stack_traj.npz")
    np.savez_compressed(path, **out)
    print(f"{path}: {len(rec['q'])} samples, contacts {out['ncon'].min()}..{out['ncon'].max()}, final reward {reward[0]}, "
          f"fp32-oracle vs fp64-oracle |dqd| spread {spread:.2e}")


if __name__ == "__main__":
    main()
