"""Trajectory store + dataset export.  CPU tier: the LeRobot-layout writer round-trips.  GPU tier: the device
recorder / compaction kernels against a numpy restatement of the reference collector (examples/franka/pick_cube_state.py:86-113),
bit-exact."""
import json
import os

import numpy as np
import pytest
import torch


def reference_collect(agent_list, env_list, action_list, reward_list):
    """The reference collector, restated (pick_cube_state.py:94-113): stack the per-step host copies into (T, B, dim) arrays and
    emit, env by env, the frames of every env whose reward was ever > 0.  Returns (episodes (k, T, F), env ids)."""
    agent, envs, act, rew = (np.stack(x) for x in (agent_list, env_list, action_list, reward_list))
    eps, ids = [], []
    for b in range(rew.shape[1]):
        if np.any(rew[:, b] > 0):
            eps.append(np.concatenate([agent[:, b], envs[:, b], act[:, b], rew[:, b, None]], axis=1).astype(np.float32))
            ids.append(b)
    T, F = rew.shape[0], agent.shape[2] + envs.shape[2] + act.shape[2] + 1
    return (np.stack(eps) if eps else np.zeros((0, T, F), np.float32)), np.array(ids, np.int32)


def test_lerobot_writer_roundtrip(tmp_path):
    import pyarrow.parquet as pq
    from gym_genesis_b200.recorder import LeRobotWriter
    rng = np.random.RandomState(0)
    w = LeRobotWriter(str(tmp_path), fps=60, robot_type="franka")
    eps = [{"observation.state": rng.rand(T, 9).astype(np.float32), "observation.environment_state": rng.rand(T, 11).astype(np.float32),
            "action": rng.rand(T, 9).astype(np.float32)} for T in (200, 200, 7)]
    for e in eps:
        w.add_episode(e, "pick cube")
    info = w.finalize()
    assert info["total_episodes"] == 3 and info["total_frames"] == 407 and info["fps"] == 60 and info["codebase_version"] == "v2.1"
    assert info["features"]["action"] == {"dtype": "float32", "shape": [9], "names": None}
    start = 0
    for k, e in enumerate(eps):
        t = pq.read_table(os.path.join(tmp_path, info["data_path"].format(episode_chunk=0, episode_index=k)))
        for name, arr in e.items():
            assert np.array_equal(np.array(t[name].to_pylist(), dtype=np.float32), arr)
        T = arr.shape[0]
        assert t["index"].to_pylist() == list(range(start, start + T)) and set(t["episode_index"].to_pylist()) == {k}
        assert np.allclose(t["timestamp"].to_pylist(), np.arange(T) / 60.0)
        start += T
    lines = [json.loads(x) for x in open(os.path.join(tmp_path, "meta", "episodes.jsonl"))]
    assert [x["length"] for x in lines] == [200, 200, 7] and lines[0]["tasks"] == ["pick cube"]
    assert json.loads(open(os.path.join(tmp_path, "meta", "tasks.jsonl")).readline()) == {"task_index": 0, "task": "pick cube"}


def test_reference_collect_selection():
    T, B = 5, 4
    rew = np.zeros((T, B), np.float32); rew[3, 2] = 1; rew[0, 0] = 1
    ag = [np.full((B, 2), t, np.float32) for t in range(T)]
    eps, ids = reference_collect(ag, ag, ag, list(rew))
    assert ids.tolist() == [0, 2] and eps.shape == (2, 5, 7) and eps[1, 3, -1] == 1


@pytest.mark.gpu
@pytest.mark.parametrize("robot,n,lift", [("franka", 1500, "some"), ("franka", 64, "none"), ("so101", 200, "all"), ("franka", 1, "all")])
def test_device_recorder_matches_reference_collector(robot, n, lift):
    from gym_genesis_b200.model import get_model
    from gym_genesis_b200.recorder import TrajectoryRecorder
    from gym_genesis_b200.sim import BatchedSim
    m = get_model(robot)
    if robot == "so101":
        from gym_genesis_b200.model_so101 import ISLAND_TOP_Z, so101_cube_pick_model
        m = so101_cube_pick_model(success_height=ISLAND_TOP_Z + 0.1)     # the literal reference threshold (0.1) is always exceeded
    sim = BatchedSim(m, n)
    T = 12
    rec = TrajectoryRecorder(sim, T)
    assert rec.frame_dim == sim.agent_dim + 11 + sim.nu + 1
    rng = np.random.RandomState(4)
    sim.reset(None, None)
    # throw a subset of the cubes into the air so that their reward becomes 1 for a few steps
    q, qd, qa = sim.get_state()
    up = {"some": torch.as_tensor(rng.rand(n) < 0.3), "none": torch.zeros(n, dtype=torch.bool), "all": torch.ones(n, dtype=torch.bool)}[lift].to(sim.device)
    q[up, sim.nu + 2] = m.success_z + 0.05
    sim.set_state(q, qd, qa)
    rec.begin_episode()
    ag, en, ac, rw = [], [], [], []
    home = torch.as_tensor(m.home_q, dtype=torch.float32, device=sim.device)
    for t in range(T):
        action = (home + 0.1 * torch.randn((n, sim.nu), device=sim.device)).contiguous()
        sim.step(action)
        rec.record(t, action)
        ag.append(sim.obs_agent.cpu().numpy()); en.append(sim.obs_env.cpu().numpy()); ac.append(action.cpu().numpy()); rw.append(sim.reward.cpu().numpy())
    eps_ref, ids_ref = reference_collect(ag, en, ac, rw)
    # the raw store, frame by frame
    traj = rec.traj.cpu().numpy()
    sl = rec.slices
    assert np.array_equal(traj[:, :, sl["agent_pos"]], np.stack(ag)) and np.array_equal(traj[:, :, sl["environment_state"]], np.stack(en))
    assert np.array_equal(traj[:, :, sl["action"]], np.stack(ac)) and np.array_equal(traj[:, :, sl["reward"]][..., 0], np.stack(rw))
    frames, env_ids = rec.export_successful()
    assert np.array_equal(env_ids.numpy(), ids_ref)
    assert frames.shape == eps_ref.shape and np.array_equal(frames.numpy(), eps_ref)
    if lift == "some":
        assert 0 < len(ids_ref) < n
    elif lift == "none":
        assert len(ids_ref) == 0
    else:
        assert len(ids_ref) == n
    # a second episode re-uses the buffers: flags cleared, nothing selected when nothing is lifted
    rec.begin_episode()
    sim.reset(None, None)
    sim.step(home.repeat(n, 1))
    rec.record(0, home.repeat(n, 1))
    _, ids2 = rec.export_successful()
    assert len(ids2) == 0
    with pytest.raises(Exception):
        rec.record(T, home.repeat(n, 1))
    sim.close()


@pytest.mark.gpu
@pytest.mark.parametrize("robot", ["franka", "so101"])
def test_device_recorder_stacking_scenes(robot):
    """The same store for the stacking tasks (gg_stack_record / gg_stack_export_successful): 14-value environment_state, agent_pos of
    9 (Franka: eef pose + fingers) or 6 (SO-101: joint positions) values."""
    from gym_genesis_b200.model import get_model
    from gym_genesis_b200.recorder import TrajectoryRecorder
    from gym_genesis_b200.sim import BatchedStackSim
    m = get_model(robot, "cube_stack")
    n, T = 300, 8
    sim = BatchedStackSim(m, n, seed=2)
    rec = TrajectoryRecorder(sim, T)
    assert rec.frame_dim == sim.agent_dim + 14 + sim.nu + 1 and sim.agent_dim == (9 if robot == "franka" else 6)
    sim.reset(None, None)
    q, qd, qa = sim.get_state()
    rng = np.random.RandomState(1)
    up = torch.as_tensor(rng.rand(n) < 0.25).to(sim.device)            # put cube_1 on top of cube_2 in a quarter of the envs
    nu = sim.nu
    q[up, nu:nu + 3] = q[up, nu + 7:nu + 10] + torch.tensor([0.0, 0.0, 0.0402], device=sim.device)
    sim.set_state(q, qd, qa)
    rec.begin_episode()
    ag, en, ac, rw = [], [], [], []
    home = torch.as_tensor(m.home_q, dtype=torch.float32, device=sim.device)
    for t in range(T):
        action = (home + 0.05 * torch.randn((n, nu), device=sim.device)).contiguous()
        sim.step(action)
        rec.record(t, action)
        ag.append(sim.obs_agent.cpu().numpy()); en.append(sim.obs_env.cpu().numpy()); ac.append(action.cpu().numpy()); rw.append(sim.reward.cpu().numpy())
    eps_ref, ids_ref = reference_collect(ag, en, ac, rw)
    frames, env_ids = rec.export_successful()
    assert np.array_equal(env_ids.numpy(), ids_ref) and 0 < len(ids_ref) < n
    assert frames.shape == eps_ref.shape and np.array_equal(frames.numpy(), eps_ref)
    assert set(ids_ref.tolist()) >= set(np.nonzero(up.cpu().numpy())[0].tolist()[:5])
    sim.close()
