"""GPU tier of the batched IK (`gg_ik`): the CUDA kernel through the C ABI against the oracle, and the
reference expert's closed loop (IK -> env.step, examples/franka/pick_cube_state.py:16-54,86-93) run entirely on the GPU."""
import numpy as np
import pytest
import torch

from gym_genesis_b200.model import FRANKA_HOME, get_model
from helpers import f32
from oracle.binding import Oracle
from test_ik_cpu import TOL_IK_Q, reachable_targets

pytestmark = pytest.mark.gpu


def _np(t):
    return t.detach().cpu().numpy().astype(np.float64)


@pytest.mark.parametrize("robot,n", [("franka", 1000), ("so101", 257), ("franka", 1)])
def test_ik_matches_oracle(robot, n):
    from gym_genesis_b200.sim import BatchedSim
    m = get_model(robot)
    orc = Oracle(m, np.float64)
    rng = np.random.RandomState(5)
    q0, tp, tq = reachable_targets(m, orc, rng, n, spread=0.15 if robot == "so101" else 0.25)
    sim = BatchedSim(m, n)
    for iters in (1, 4, 20):
        q, err, it = sim.inverse_kinematics(pos=f32(tp), quat=f32(tq), init_qpos=f32(q0), max_solver_iters=iters, pos_tol=0.0, rot_tol=0.0,
                                            return_error=True, return_iters=True)
        qr, er, ir = orc.ik(q0, tp, tq, max_iters=iters, pos_tol=0.0, rot_tol=0.0)
        assert (it.cpu().numpy() == iters).all()
        assert np.abs(_np(q) - qr).max() < TOL_IK_Q, (iters, np.abs(_np(q) - qr).max())
        assert np.abs(_np(err) - er).max() < 1e-4
    q, err, it = sim.inverse_kinematics(pos=f32(tp), quat=f32(tq), init_qpos=f32(q0), return_error=True, return_iters=True)
    qr, er, ir = orc.ik(q0, tp, tq)
    same = it.cpu().numpy() == ir
    assert same.mean() > 0.9
    assert np.abs(_np(q)[same] - qr[same]).max() < TOL_IK_Q
    if robot == "franka":
        e = _np(err)
        assert (e[:, 0] < 5e-4).all() and (e[:, 1] < 5e-3).all()
    assert (_np(q)[:, m.n_arm:] == f32(q0)[:, m.n_arm:].astype(np.float64)).all()
    sim.close()


def test_ik_from_current_state_envs_idx_and_masks():
    from gym_genesis_b200.sim import BatchedSim
    m = get_model("franka")
    orc = Oracle(m, np.float64)
    n = 96
    sim = BatchedSim(m, n)
    rng = np.random.RandomState(1)
    xy = np.stack([rng.uniform(0.45, 0.8, n), rng.uniform(-0.25, 0.25, n)], 1).astype(np.float32)
    sim.reset(None, xy)
    q_now = _np(sim.get_state()[0])[:, :9]
    cube = _np(sim.obs_env)[:, :3]
    tp = cube + np.array([0.0, 0.0, 0.115])
    tq = np.tile([0.0, 1.0, 0.0, 0.0], (n, 1))
    # init_qpos=None: start from the current joint positions of the env
    q = _np(sim.inverse_kinematics(pos=f32(tp), quat=f32(tq), max_solver_iters=6, pos_tol=0.0, rot_tol=0.0))
    qr, _, _ = orc.ik(q_now, f32(tp).astype(np.float64), tq, max_iters=6, pos_tol=0.0, rot_tol=0.0)
    assert np.abs(q - qr).max() < TOL_IK_Q
    # subset of envs, in scrambled order
    idx = rng.permutation(n)[:37]
    qs = _np(sim.inverse_kinematics(pos=f32(tp[idx]), quat=f32(tq[idx]), envs_idx=torch.as_tensor(idx), max_solver_iters=6, pos_tol=0.0, rot_tol=0.0))
    assert np.abs(qs - qr[idx]).max() < TOL_IK_Q
    # position only / z only
    qp, ep = sim.inverse_kinematics(pos=f32(tp), return_error=True)
    qpr, epr, _ = orc.ik(q_now, f32(tp).astype(np.float64), None)
    ep = _np(ep)
    solved = epr[:, 0] < 4e-4            # cubes at the far corners of the spawn area are out of reach: those rows stay unsolved
    assert (ep[:, 1] == 0).all() and np.abs(ep[solved, 0] - epr[solved, 0]).max() < 1e-4
    assert solved.mean() > 0.8 and (ep[solved, 0] < 5e-4).all() and (ep[~solved, 0] > 4e-4).all()
    qz, ez = sim.inverse_kinematics(pos=f32(tp), pos_mask=(False, False, True), return_error=True)
    qzr, _, _ = orc.ik(q_now, f32(tp).astype(np.float64), None, mask6=[0, 0, 1, 0, 0, 0])
    assert np.abs(_np(qz) - qzr).max() < TOL_IK_Q
    # the simulation state is untouched
    assert np.array_equal(_np(sim.get_state()[0])[:, :9], q_now)
    # error paths
    with pytest.raises(ValueError):
        sim.inverse_kinematics()
    with pytest.raises(ValueError):
        sim.inverse_kinematics(pos=f32(tp[:5]), quat=f32(tq[:6]))
    sim.close()


def test_expert_closed_loop_on_gpu_matches_oracle_and_succeeds():
    """The reference expert (hover / stabilize / grasp / grasp / lift x 40 steps; IK target = cube + offset, hand down; fingers
    0.04 / -0.02) through the gymnasium surface: env.get_robot().inverse_kinematics(...) + env.step(...), nothing on the host.
    Free-space stages must track the fp64 oracle closed loop; overall the expert must lift a fair share of the cubes."""
    import gym_genesis_b200  # noqa: F401  (registers the ids)
    from gym_genesis_b200.gym_api import gym
    n = 64
    env = gym.make("gym_genesis/CubePick-v0", robot="franka", num_envs=n, enable_pixels=False).unwrapped
    obs, _ = env.reset(seed=7)
    task = env._env if hasattr(env, "_env") else env.task
    sim = task.sim
    m = get_model("franka")
    orc = Oracle(m, np.float64)
    Q, QD, QA = (_np(t) for t in sim.get_state())
    robot = env.get_robot()
    eef = robot.get_link("hand")
    quat = torch.tensor([0.0, 1.0, 0.0, 0.0], device=sim.device).expand(n, -1)
    k_ref = 4                                   # envs also advanced by the oracle closed loop
    Qo, QDo, QAo = Q[:k_ref].copy(), QD[:k_ref].copy(), QA[:k_ref].copy()
    success = torch.zeros(n, dtype=torch.bool, device=sim.device)
    step = 0
    for dz, grip in [(0.115, 0.04), (0.115, 0.04), (0.03, -0.02), (0.03, -0.02), (0.25, -0.02)]:
        for _ in range(40):
            cube = obs["environment_state"][:, :3]
            target = cube + torch.tensor([0.0, 0.0, dz], device=sim.device)
            qpos = robot.inverse_kinematics(link=eef, pos=target, quat=quat, envs_idx=torch.arange(n, device=sim.device))
            action = torch.cat([qpos[:, :-2], torch.full((n, 2), grip, device=sim.device)], dim=1)
            if step < 80:
                # oracle closed loop on the first k_ref envs (free-space stages only: contact-rich phases are chaotic)
                tpo = Qo[:, 9:12] + np.array([0.0, 0.0, dz])
                qo, _, _ = orc.ik(Qo[:, :9], tpo, np.tile([0.0, 1.0, 0.0, 0.0], (k_ref, 1)))
                ao = np.concatenate([qo[:, :7], np.full((k_ref, 2), grip)], 1)
                assert np.abs(_np(action[:k_ref]) - ao).max() < 2e-3, (step, np.abs(_np(action[:k_ref]) - ao).max())
                orc.step(Qo, QDo, QAo, ao, iterations=100, tol=1e-10)
            obs, reward, terminated, truncated, info = env.step(action)
            success |= terminated
            if step < 80 and step % 10 == 9:
                qg = _np(sim.get_state()[0])[:k_ref]
                assert np.abs(qg[:, :9] - Qo[:, :9]).max() < 5e-3, (step, np.abs(qg[:, :9] - Qo[:, :9]).max())
            step += 1
    rate = success.float().mean().item()
    # the same closed loop in the fp64 oracle lifts ~60 % of the cubes (the open-loop expert grasps blindly at the 120th step)
    assert rate > 0.35, rate
    env.close()


def test_ik_resampling_kernel_matches_sequential_oracle():
    """gg_ik with max_samples: one warp per unsolved row runs the samples as its 32 lanes and keeps the first success in sample order --
    the result of the oracle's sequential loop over the same Philox draws."""
    from test_ik_cpu import bad_start_problem
    from gym_genesis_b200.sim import BatchedSim
    m = get_model("franka")
    orc = Oracle(m, np.float64)
    n = 512
    q0, tp, tq = bad_start_problem(m, orc, np.random.RandomState(5), n)
    seed = (3 << 32) | 11
    sim = BatchedSim(m, n, seed=seed)
    qs, es, its = sim.inverse_kinematics(pos=f32(tp), quat=f32(tq), init_qpos=f32(q0), return_error=True, return_iters=True)
    ok_first = ((es[:, 0] < 5e-4) & (es[:, 1] < 5e-3)).cpu().numpy()
    assert 0.2 < ok_first.mean() < 0.7
    q, e, it = sim.inverse_kinematics(pos=f32(tp), quat=f32(tq), init_qpos=f32(q0), return_error=True, return_iters=True, max_samples=50)
    q, e, it = q.cpu().numpy().astype(np.float64), e.cpu().numpy(), it.cpu().numpy()
    ok = (e[:, 0] < 5e-4) & (e[:, 1] < 5e-3)
    assert ok.mean() > 0.99                                         # Genesis' default of 50 attempts solves every reachable target
    assert np.array_equal(q[ok_first], qs.cpu().numpy().astype(np.float64)[ok_first])     # solved rows are not touched by the second pass
    qr, er, ir = orc.ik(q0, tp, tq, max_samples=50, seed=seed)
    # forward kinematics of the kernel's answers reach the targets
    for i in np.nonzero(ok)[0][:32]:
        xp, _ = orc.fk(m.full_q(q[i]))
        assert np.abs(xp[m.eef_body] - tp[i]).max() < 6e-4
    # same chosen sample wherever fp32 and fp64 agree on which attempt succeeded first (the lanes of a round all run to the end, so the
    # kernel's update count is not comparable; compare the joint vectors)
    close = np.abs(q - qr).max(axis=1) < 1e-3
    assert close.mean() > 0.85, close.mean()
    # deterministic
    q2 = sim.inverse_kinematics(pos=f32(tp), quat=f32(tq), init_qpos=f32(q0), max_samples=50)
    assert torch.equal(q2, torch.as_tensor(q, dtype=torch.float32, device=q2.device))
    # out-of-range env indices are rejected on the host
    with pytest.raises(ValueError):
        sim.inverse_kinematics(pos=f32(tp[:4]), quat=f32(tq[:4]), envs_idx=[0, 1, 2, n])
    with pytest.raises(ValueError):
        sim.inverse_kinematics(pos=f32(tp[:4]), quat=f32(tq[:4]), envs_idx=[-1, 1, 2, 3])
    sim.close()
