"""State-only expert data collection for CubePick-v0 (Franka) on the B200 backend.

Same program as the reference's `examples/franka/pick_cube_state.py` -- gym.make, the hover / stabilize / grasp / grasp / lift
expert driven by `robot.inverse_kinematics`, save the envs whose reward was ever > 0 as LeRobot episodes -- with the three
host round trips per step removed: IK runs on the GPU (`gg_ik`), frames are recorded on the GPU (`gg_record`) and the successful
envs come back with one copy per episode (`gg_export_successful`).

    python examples/franka/pick_cube_state.py --num-envs 4096 --episodes 2 --out data/cube_genesis
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import gym_genesis  # noqa: F401,E402  (registers gym_genesis/CubePick-v0 on the B200 backend)
from gym_genesis_b200.gym_api import gym  # noqa: E402  (gymnasium when installed)
from gym_genesis_b200.recorder import LeRobotWriter, TrajectoryRecorder, write_lerobot_dataset  # noqa: E402


def expert_policy(robot, observation, stage):
    """examples/franka/pick_cube_state.py:16-54 of the reference, unchanged in behaviour."""
    cube_pos = observation["environment_state"][:, :3]
    B, device = cube_pos.shape[0], cube_pos.device
    quat = torch.tensor([0, 1, 0, 0], dtype=torch.float32, device=device).expand(B, -1)
    eef = robot.get_link("hand")
    dz, grip = {"hover": (0.115, 0.04), "stabilize": (0.115, 0.04), "grasp": (0.03, -0.02), "lift": (0.25, -0.02)}[stage]
    target_pos = cube_pos + torch.tensor([0.0, 0.0, dz], device=device)
    qpos = robot.inverse_kinematics(link=eef, pos=target_pos, quat=quat, envs_idx=torch.arange(B, device=device))
    return torch.cat([qpos[:, :-2], torch.full((B, 2), grip, device=device)], dim=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--num-envs", type=int, default=4096)
    ap.add_argument("--episodes", type=int, default=2)
    ap.add_argument("--out", default="data/cube_genesis")
    args = ap.parse_args()
    env = gym.make("gym_genesis/CubePick-v0", enable_pixels=False, robot="franka", num_envs=args.num_envs).unwrapped
    sim = env._env.sim
    rec = TrajectoryRecorder(sim, horizon=200)
    writer = LeRobotWriter(args.out, fps=60, robot_type="franka")
    for ep in range(args.episodes):
        t0 = time.perf_counter()
        obs, _ = env.reset()
        rec.begin_episode()
        t = 0
        for stage in ["hover", "stabilize", "grasp", "grasp", "lift"]:
            for _ in range(40):
                action = expert_policy(env.get_robot(), obs, stage)
                obs, reward, done, _, info = env.step(action)
                rec.record(t, action)
                t += 1
        frames, env_ids = rec.export_successful()
        t1 = time.perf_counter()
        write_lerobot_dataset(writer, rec, frames, "pick cube")
        print(f"episode {ep}: {frames.shape[0]} / {args.num_envs} envs lifted the cube; simulate + record + export {t1 - t0:.2f} s "
              f"({args.num_envs * 200 / (t1 - t0) / 1e6:.2f} M env-steps/s), parquet {time.perf_counter() - t1:.2f} s")
    info = writer.finalize()
    print(f"wrote {info['total_episodes']} episodes / {info['total_frames']} frames to {args.out}")
    env.close()


if __name__ == "__main__":
    main()
