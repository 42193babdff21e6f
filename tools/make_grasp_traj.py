"""Generates tests/golden/grasp_traj.npz: the reference expert's stage machine
(examples/franka/pick_cube_state.py:16-54,86-93 -- hover / stabilize / grasp / grasp / lift, 40 steps each, hand target =
cube + offset with quat [0,1,0,0], fingers 0.04 open / -0.02 closed) run CLOSED-LOOP in the fp64 oracle for K seeded cube
positions, with a numpy damped-least-squares IK standing in for robot.inverse_kinematics.  The recorded joint-target
sequences are replayed OPEN-LOOP by the GPU tests and by bench.py's contact-rich workload B."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from gym_genesis_b200.model import FRANKA_HOME, franka_cube_pick_model, quat_to_mat  # noqa: E402
from helpers import ik_numpy  # noqa: E402
from oracle.binding import Oracle  # noqa: E402

K = int(sys.argv[1]) if len(sys.argv) > 1 else 32
m = franka_cube_pick_model()
o = Oracle(m, np.float64)
rng = np.random.RandomState(2024)
xy = np.stack([rng.uniform(0.45, 0.80, K), rng.uniform(-0.25, 0.25, K)], 1).astype(np.float32)
Rdown = quat_to_mat([0, 1, 0, 0])
actions = np.zeros((K, 200, 9), np.float32)
cube = np.zeros((K, 200, 3), np.float32)
qrob = np.zeros((K, 200, 9), np.float32)
ncon = np.zeros((K, 200), np.int32)
reward = np.zeros((K, 200), np.float32)
for k in range(K):
    Q, QD, QA = o.new_state(1)
    o.reset(Q, QD, QA, xy[k:k + 1].astype(np.float64))
    o.step(Q, QD, QA, FRANKA_HOME[None].copy())                 # settle step of reset()
    t = 0
    for dz, grip in [(0.115, 0.04), (0.115, 0.04), (0.03, -0.02), (0.03, -0.02), (0.25, -0.02)]:
        for _ in range(40):
            qt = ik_numpy(m, Q[0, 9:12] + np.array([0, 0, dz]), Rdown, Q[0, :9].copy())
            a = np.concatenate([qt[:7], [grip, grip]]).astype(np.float32)
            actions[k, t] = a
            info, _ = o.step(Q, QD, QA, a[None].astype(np.float64), iterations=100, tol=1e-10)
            cube[k, t] = Q[0, 9:12]; qrob[k, t] = Q[0, :9]; ncon[k, t] = info[0, 0]
            reward[k, t] = o.obs(Q)[2][0]
            t += 1
    print(k, xy[k], "final cube", cube[k, -1], "success", reward[k].max(), "max contacts", ncon[k].max(), flush=True)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "grasp_traj.npz"), xy=xy, actions=actions, cube=cube, qrob=qrob,
                    ncon=ncon, reward=reward)
print("success rate", (reward.max(1) > 0).mean())
