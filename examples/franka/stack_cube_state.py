"""State-only expert data collection for CubeStack-v0 (Franka) on the B200 backend.

Same program as the reference's `examples/franka/stack_cube_state.py` -- gym.make, the hover / grasp / lift / place / release expert
(Cartesian way-points -> `robot.inverse_kinematics` -> joint-space interpolation, 80 actions per stage), save the envs whose reward
was ever > 0 as LeRobot episodes -- with the host round trips removed: IK runs on the GPU (`gg_ik`), frames are recorded on the
GPU (`gg_stack_record`) and the successful envs come back with one copy per episode (`gg_stack_export_successful`).
`--scene kitchen` (default) is what the reference's `CubeStack-v0` + franka builds -- the Panda at morph scale 0.6 next to the kitchen
island -- and uses the reference's way-point heights literally (`:33-49`: hover 0.25, grasp 0.045, lift 0.28, place / release 0.18
above the cube centres, tuned for that 0.6-scale gripper); `--scene empty` is the full-size Panda on the ground plane
(`FrankaCubeStackEmptyBatch`) with the heights shifted by its hand-to-pad distance (grasp 0.105, place / release 0.148).

    python examples/franka/stack_cube_state.py --num-envs 1024 --episodes 1 --out data/stack_cube
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import gym_genesis  # noqa: F401,E402  (registers gym_genesis/CubeStack-v0 on the B200 backend)
from gym_genesis_b200.gym_api import gym  # noqa: E402  (gymnasium when installed)
from gym_genesis_b200.recorder import LeRobotWriter, TrajectoryRecorder, write_lerobot_dataset  # noqa: E402

STAGES = ["hover", "grasp", "lift", "place", "release"]
GRIP_OPEN, GRIP_CLOSED = 0.04, -0.02
# way-points per stage (8 for the straight moves, 4 + 4 + 3 via the hover point) -> 80 // (n - 1) actions per segment (:84-91)
N_WP = {"hover": 8, "grasp": 8, "lift": 8, "place": 11, "release": 11}
HORIZON = sum((80 // (N_WP[s] - 1)) * (N_WP[s] - 1) for s in STAGES)


# hand height above the cube centre per stage: (hover, grasp, lift, carry, place / release)
HEIGHTS = {"kitchen": (0.25, 0.045, 0.28, 0.25, 0.18), "empty": (0.25, 0.105, 0.28, 0.25, 0.148)}


def expert_policy(robot, obs, stage, heights=HEIGHTS["empty"], grip_open=GRIP_OPEN):
    """Reference `expert_policy` (stack_cube_state.py:18-109): returns a list of (B, 9) joint-target tensors for one stage."""
    es = obs["environment_state"]
    B, dev = es.shape[0], es.device
    eef = robot.get_link("hand")
    quat = torch.tensor([0.0, 1.0, 0.0, 0.0], device=dev).repeat(B, 1)
    cube1, cube2 = es[:, :3], es[:, 11:14]
    off = lambda z, y=0.0: torch.tensor([0.0, y, z], device=dev)
    hover_pos = None
    h_hover, h_grasp, h_lift, h_carry, h_place = heights
    if stage == "hover":
        target, grip = cube1 + off(h_hover), grip_open
    elif stage == "grasp":
        target, grip = cube1 + off(h_grasp), GRIP_CLOSED
    elif stage == "lift":
        target, grip = cube1 + off(h_lift), GRIP_CLOSED
    elif stage == "place":
        hover_pos, target, grip = cube2 + off(h_carry), cube2 + off(h_place, 0.004), GRIP_CLOSED      # :44: 4 mm y offset
    elif stage == "release":
        hover_pos, target, grip = cube2 + off(h_carry), cube2 + off(h_place), grip_open
    else:
        raise ValueError(f"Unknown stage: {stage}")
    current = eef.get_pos(envs_idx=torch.arange(B, device=dev))
    wps = []
    if hover_pos is not None:
        wps += [(1 - a) * current + a * hover_pos for a in torch.linspace(0, 1, 4)]
        wps += [(1 - a) * hover_pos + a * target for a in torch.linspace(0, 1, 4)]
        wps += [target] * 3
    else:
        wps += [(1 - a) * current + a * target for a in torch.linspace(0, 1, 8)]
    init_q = robot.get_qpos(envs_idx=torch.arange(B, device=dev))
    q_wps = []
    for wp in wps:
        init_q = robot.inverse_kinematics(link=eef, pos=wp, quat=quat, init_qpos=init_q)      # continuity: start from the last solution
        q_wps.append(init_q)
    per = 80 // (len(q_wps) - 1)
    path = []
    for i in range(len(q_wps) - 1):
        for t in range(per):
            a = t / (per - 1)
            path.append(((1 - a) * q_wps[i] + a * q_wps[i + 1]).clone())
    if stage == "grasp":                                       # close over the last five actions (:96-103)
        for i, q in enumerate(path):
            k = i - (len(path) - 5)
            q[:, -2:] = grip_open if k < 0 else (1 - k / 5) * grip_open + (k / 5) * GRIP_CLOSED
    else:
        for q in path:
            q[:, -2:] = grip
    return path


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--num-envs", type=int, default=1024)
    ap.add_argument("--episodes", type=int, default=1)
    ap.add_argument("--out", default="data/stack_cube")
    ap.add_argument("--scene", default="kitchen", choices=["kitchen", "empty"])
    args = ap.parse_args()
    env = gym.make("gym_genesis/CubeStack-v0", enable_pixels=False, robot="franka", num_envs=args.num_envs, scene=args.scene).unwrapped
    sim = env._env.sim
    rec = TrajectoryRecorder(sim, horizon=HORIZON)
    writer = LeRobotWriter(args.out, fps=60, robot_type="franka")
    for ep in range(args.episodes):
        t0 = time.perf_counter()
        obs, _ = env.reset()
        t = 0
        for stage in STAGES:
            path = expert_policy(env.get_robot(), obs, stage, HEIGHTS[args.scene])
            if t == 0:
                rec.begin_episode()
            for action in path:
                obs, reward, done, _, info = env.step(action)
                rec.record(t, action)
                t += 1
        stacked = int(reward.sum().item())
        frames, env_ids = rec.export_successful()
        t1 = time.perf_counter()
        write_lerobot_dataset(writer, rec, frames, "pick up the red cube and place it on top of the green cube")
        print(f"episode {ep}: {frames.shape[0]} / {args.num_envs} envs saw reward > 0, {stacked} end stacked; simulate + IK + record + export "
              f"{t1 - t0:.2f} s ({args.num_envs * t / (t1 - t0) / 1e6:.2f} M env-steps/s), parquet {time.perf_counter() - t1:.2f} s")
    info = writer.finalize()
    print(f"wrote {info['total_episodes']} episodes / {info['total_frames']} frames to {args.out}")
    env.close()


if __name__ == "__main__":
    main()
