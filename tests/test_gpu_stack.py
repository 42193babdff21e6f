"""GPU tier of the multi-object (CubeStack) path, SURVEY.md 8f-1: the CUDA kernels called through the C ABI
(BatchedStackSim -> gg_stack_* in libgg_b200.so) against the oracle and the committed golden trajectory."""
import os

import numpy as np
import pytest
import torch

from gym_genesis_b200.model import FRANKA_HOME, franka_cube_stack_model
from helpers import (ROLLOUT_POS_TOL, TOL_Q_ABS, f32, stack_cluster_states, stack_clutter_states, stack_pile_states,
                     stack_random_states, stack_substep_errors)

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stack_traj.npz")


@pytest.fixture(scope="module")
def sm():
    return franka_cube_stack_model()


@pytest.fixture(scope="module")
def orc(sm):
    from oracle.binding import Oracle, build
    build()
    return Oracle(sm, np.float64)


@pytest.fixture(scope="module")
def Sim():
    from gym_genesis_b200.sim import BatchedStackSim
    return BatchedStackSim


def _np(t):
    return t.detach().cpu().numpy().astype(np.float64)


def _asm(o, Q, QD, T):
    return np.array([np.abs(o.dynamics(Q[k], QD[k], T[k])["a_smooth"]).max() * 0.01 for k in range(Q.shape[0])])


def _one_substep(sim, Q, QD, QA, T):
    sim.set_state(Q, QD, QA)
    sim.step(torch.as_tensor(f32(T)))
    q, qd, qa = sim.get_state()
    return _np(q), _np(qd), _np(qa)


def test_state_roundtrip_and_layout(Sim, sm):
    n = 37
    sim = Sim(sm, n)
    assert sim.state.shape == (3 * 9 + 19 * 5, 64)
    rng = np.random.default_rng(0)
    Q, QD, QA = f32(rng.normal(size=(n, 44))), f32(rng.normal(size=(n, 39))), f32(rng.normal(size=(n, 39)))
    sim.set_state(Q, QD, QA)
    q, qd, qa = sim.get_state()
    assert (q.cpu().numpy() == Q).all() and (qd.cpu().numpy() == QD).all() and (qa.cpu().numpy() == QA).all()
    st = sim.state.cpu().numpy()
    assert (st[0:9, :n].T == Q[:, :9]).all() and (st[18 + 13:18 + 13 + 7, :n].T == Q[:, 16:23]).all()


def test_single_substep_random_states(Sim, sm, orc):
    n = 192
    rng = np.random.default_rng(31)
    Q, QD, T = stack_random_states(sm, rng, n)
    QA = np.zeros_like(QD)
    sim = Sim(sm, n)
    q, qd, qa = _one_substep(sim, Q, QD, QA, T)
    qo, qdo, qao = Q.copy(), QD.copy(), QA.copy()
    info, _ = orc.step(qo, qdo, qao, T, iterations=100, tol=1e-12)
    nc, it = sim.step_info()
    assert (nc.cpu().numpy() == info[:, 0]).all()             # identical contact counts, cube-cube pairs included
    eq, ev = stack_substep_errors(Q, QD, qo, qdo, q, qd, _asm(orc, Q, QD, T))
    assert eq.max() < TOL_Q_ABS and ev.max() < 1.0, (eq.max(), ev.max(), int(np.argmax(ev)))
    # observations / reward / flags against the oracle on the GPU's own state: flags bit-exact
    a, e, r, t = orc.obs(q)
    assert np.abs(_np(sim.obs_agent) - a).max() < 2e-6 and np.abs(_np(sim.obs_env) - e).max() < 2e-6
    assert (sim.terminated.cpu().numpy().astype(bool) == t).all() and t[2]
    assert (sim.reward.cpu().numpy() == r.astype(np.float32)).all()
    assert not sim.truncated.any()


def test_single_substep_golden_stack_trajectory(Sim, sm):
    """Every sample of the committed pick-and-stack rollout (free space, grasp with 36-52 contacts, carry, cube_1 lowered onto
    cube_2, release) against the oracle's stored next state."""
    g = np.load(GOLDEN)
    Q, QD, QA, T = g["q"], g["qd"], g["qacc"], g["act"]
    sim = Sim(sm, Q.shape[0])
    q, qd, qa = _one_substep(sim, Q, QD, QA, T)
    nc, it = sim.step_info()
    assert (nc.cpu().numpy() == g["ncon"]).all()
    eq, ev = stack_substep_errors(Q, QD, g["qn"], g["qdn"], q, qd, g["asm"])
    assert eq.max() < TOL_Q_ABS and ev.max() < 1.0, (eq.max(), ev.max(), int(np.argmax(ev)))
    assert it.max().item() <= 10


def test_all_bodies_active_two_rows_per_lane(Sim, sm, orc):
    """Robot and all five cubes within reach of each other: the cooperative solve runs on all 39 dofs (two rows per lane)."""
    Q, QD, QA, T = stack_cluster_states(np.load(GOLDEN))
    Q, QD, QA, T = (np.tile(x, (8, 1)) for x in (Q, QD, QA, T))          # a full warp of envs for the thread-per-env kernels
    sim = Sim(sm, Q.shape[0])
    q, qd, qa = _one_substep(sim, Q, QD, QA, T)
    qo, qdo, qao = Q.copy(), QD.copy(), QA.copy()
    info, _ = orc.step(qo, qdo, qao, T, iterations=200, tol=1e-12)
    assert (sim.step_info()[0].cpu().numpy() == info[:, 0]).all() and info[:, 0].max() >= 48
    eq, ev = stack_substep_errors(Q, QD, qo, qdo, q, qd, _asm(orc, Q, QD, T))
    assert eq.max() < TOL_Q_ABS and ev.max() < 1.0, (eq.max(), ev.max())


def test_tumbling_pile_states(Sim, sm, orc):
    """Cubes dropped on one spot with random orientations and spins (oracle rollout, sampled every 15 steps): edge-edge and
    vertex-face cube-cube contacts, cubes resting on two others, corner bounces."""
    Q, QD, QA, T = stack_pile_states(sm, orc, n=32, steps=120)
    sim = Sim(sm, Q.shape[0])
    q, qd, qa = _one_substep(sim, Q, QD, QA, T)
    qo, qdo, qao = Q.copy(), QD.copy(), QA.copy()
    info, _ = orc.step(qo, qdo, qao, T, iterations=200, tol=1e-12)
    assert (sim.step_info()[0].cpu().numpy() == info[:, 0]).all() and info[:, 0].max() >= 16
    eq, ev = stack_substep_errors(Q, QD, qo, qdo, q, qd, _asm(orc, Q, QD, T))
    assert eq.max() < TOL_Q_ABS and ev.max() < 1.0, (eq.max(), ev.max(), int(np.argmax(ev)))


def test_gripper_cube_clutter_states(Sim, sm, orc):
    """512 states with cubes overlapping the fingers, each other and the plane at random poses (up to ~75 contacts)."""
    Q, QD, T = stack_clutter_states(sm, Sim(sm, 1).gg_model, np.random.default_rng(321), 512)
    QA = np.zeros_like(QD)
    sim = Sim(sm, Q.shape[0])
    q, qd, qa = _one_substep(sim, Q, QD, QA, T)
    qo, qdo, qao = Q.copy(), QD.copy(), QA.copy()
    info, _ = orc.step(qo, qdo, qao, T, iterations=300, tol=1e-12)
    assert (sim.step_info()[0].cpu().numpy() == info[:, 0]).all() and info[:, 0].max() >= 60
    eq, ev = stack_substep_errors(Q, QD, qo, qdo, q, qd, _asm(orc, Q, QD, T))
    assert eq.max() < TOL_Q_ABS and ev.max() < 1.0, (eq.max(), ev.max(), int(np.argmax(ev)))


def test_reset_matches_reference_semantics(Sim, sm, orc):
    """reset = per-cube spawn (host stream), home pose, hold-home targets, ONE settle step, obs (cube_stack.py:119-170)."""
    n = 64
    rng = np.random.RandomState(5)
    xy = np.zeros((n, 5, 2), np.float32)
    for k, (xlo, xhi, ylo, yhi) in enumerate(sm.obj_spawn):
        xy[:, k, 0] = rng.uniform(xlo, xhi, n)
        xy[:, k, 1] = rng.uniform(ylo, yhi, n)
    sim = Sim(sm, n)
    sim.reset(None, xy)
    q, qd, qa = (_np(t) for t in sim.get_state())
    qo, qdo, qao = orc.new_state(n)
    orc.reset(qo, qdo, qao, xy.astype(np.float64))
    q0, qd0 = qo.copy(), qdo.copy()
    tgt = np.tile(FRANKA_HOME, (n, 1))
    orc.step(qo, qdo, qao, tgt, iterations=100, tol=1e-12)
    eq, ev = stack_substep_errors(q0, qd0, qo, qdo, q, qd, _asm(orc, q0, qd0, tgt))
    assert eq.max() < TOL_Q_ABS and ev.max() < 1.0, (eq.max(), ev.max())
    a, e, r, t = orc.obs(q)
    assert np.abs(_np(sim.obs_agent) - a).max() < 2e-6 and np.abs(_np(sim.obs_env) - e).max() < 2e-6
    assert (sim.episode.cpu().numpy() == 1).all() and (sim.step_count.cpu().numpy() == 0).all()


def test_masked_reset_counters_and_philox_spawn_bit_exact(Sim, sm, orc):
    n = 96
    seed = 1234
    sim = Sim(sm, n, seed=seed, first_env_id=1000)
    sim.reset(None, None)                                     # device Philox draws for every env, episode 0
    q0 = sim.get_state()[0].cpu().numpy()
    sim.configure(solver_iterations=10)
    tgt = torch.as_tensor(f32(np.tile(FRANKA_HOME, (n, 1))))
    for _ in range(3):
        sim.step(tgt)
    assert (sim.step_count.cpu().numpy() == 3).all()
    mask = np.zeros(n, bool)
    mask[::3] = True
    before = sim.get_state()[0].cpu().numpy()
    sim.step(tgt, reset_mask=mask)
    q1 = sim.get_state()[0].cpu().numpy()
    sc, ep = sim.step_count.cpu().numpy(), sim.episode.cpu().numpy()
    assert (sc[mask] == 0).all() and (sc[~mask] == 4).all() and (ep[mask] == 2).all() and (ep[~mask] == 1).all()
    # spawn xy of reset envs: Philox draws keyed by (seed, GLOBAL env id, episode, object), bit-exact for every cube that
    # lands clear of the others (an overlapping spawn is pushed apart by the settle step)
    checked = 0
    for e in np.nonzero(mask)[0]:
        ref = np.array([orc.spawn_obj(seed, 1000 + int(e), 1, k) for k in range(5)])
        for k in range(5):
            clear = all(np.abs(ref[k] - ref[j]).max() > 0.08 for j in range(5) if j != k)
            if clear:
                assert (q1[e, 9 + 7 * k: 9 + 7 * k + 2] == ref[k]).all()
                checked += 1
    assert checked > 50
    e = int(np.nonzero(mask)[0][0])
    first = np.array([orc.spawn_obj(seed, 1000 + e, 0, k) for k in range(5)])
    for k in range(5):
        if all(np.abs(first[k] - first[j]).max() > 0.08 for j in range(5) if j != k):
            assert (q0[e, 9 + 7 * k: 9 + 7 * k + 2] == first[k]).all()
    # envs outside the mask were stepped, not touched by the reset
    assert np.abs(q1[~mask, 9:] - before[~mask, 9:]).max() < 2e-2    # one 0.01 s step (overlapping spawns still separating)


def test_rollout_100_steps(Sim, sm, orc):
    """100 control steps, smooth random-walk targets, cubes resting (some pushed apart from overlapping spawns)."""
    n = 24
    rng = np.random.default_rng(21)
    sim = Sim(sm, n)
    xy = np.zeros((n, 5, 2), np.float32)
    for k, (xlo, xhi, ylo, yhi) in enumerate(sm.obj_spawn):
        xy[:, k, 0] = rng.uniform(xlo, xhi, n)
        xy[:, k, 1] = rng.uniform(ylo, yhi, n)
    sim.reset(None, xy)
    q, qd, qa = (_np(t) for t in sim.get_state())
    tgt = np.tile(FRANKA_HOME, (n, 1))
    for _ in range(100):
        tgt = tgt + rng.normal(0, 0.02, tgt.shape)
        tgt[:, 7:] = np.clip(tgt[:, 7:], -0.02, 0.04)
        t32 = f32(tgt)
        sim.step(torch.as_tensor(t32))
        orc.step(q, qd, qa, t32.astype(np.float64), iterations=100, tol=1e-10)
    qg = _np(sim.get_state()[0])
    pos_cols = list(range(9)) + [9 + 7 * k + c for k in range(5) for c in range(3)]
    err = np.abs(qg[:, pos_cols] - q[:, pos_cols]).max()
    assert err < ROLLOUT_POS_TOL, err


def test_full_size_invariants(Sim, sm):
    """16384 envs, 40 steps with random targets, Bernoulli resets and auto-reset: finite state, unit quaternions, cubes on or
    above the plane, joints inside their (soft) limits, counters consistent with the statistics."""
    n = 16384
    sim = Sim(sm, n, auto_reset=True, reset_prob=0.01, seed=3)
    sim.reset(None, None)
    g = torch.Generator(device="cuda").manual_seed(0)
    home = torch.as_tensor(f32(FRANKA_HOME), device="cuda")
    sim.stats(clear=True)
    for _ in range(40):
        a = home + (torch.rand((n, 9), device="cuda", generator=g) - 0.5) * 0.6
        a[:, 7:] = torch.rand((n, 2), device="cuda", generator=g) * 0.06 - 0.02
        sim.step(a)
    q, qd, qa = sim.get_state()
    assert torch.isfinite(q).all() and torch.isfinite(qd).all() and torch.isfinite(qa).all()
    cubes = q[:, 9:].reshape(n, 5, 7)
    assert (cubes[:, :, 3:].norm(dim=2) - 1).abs().max() < 1e-5
    assert cubes[:, :, 2].min() > 0.015 and cubes[:, :, 2].max() < 0.2
    lo = torch.as_tensor(list(sim.gg_model.lim_lo), device="cuda")
    hi = torch.as_tensor(list(sim.gg_model.lim_hi), device="cuda")
    assert (q[:, :9] > lo - 0.05).all() and (q[:, :9] < hi + 0.05).all()
    st = sim.stats().cpu().numpy()
    assert st[5] == 0 and 0.97 * 40 * n < st[0] <= 40 * n and st[7] / st[0] >= 19.0     # >= 5 x 4 plane contacts per env-step


def test_gym_surface_cube_stack(sm):
    import gym_genesis  # noqa: F401  (registers the ids on the B200 backend)
    from gym_genesis_b200.gym_api import gym
    n = 48
    env = gym.make("gym_genesis/CubeStack-v0", robot="franka", num_envs=n, scene="empty")
    u = env.unwrapped
    assert u.observation_space["agent_pos"].shape == (9,) and u.observation_space["environment_state"].shape == (14,)
    obs, info = env.reset(seed=3)
    assert obs["agent_pos"].shape == (n, 9) and obs["environment_state"].shape == (n, 14) and obs["agent_pos"].is_cuda
    first = obs["environment_state"].clone()
    a = torch.as_tensor(f32(np.tile(FRANKA_HOME, (n, 1))), device="cuda")
    obs, reward, terminated, truncated, info = u.step(a)
    assert reward.shape == (n,) and terminated.dtype == torch.bool and truncated.dtype == torch.bool and not truncated.any()
    assert (terminated == (reward == 1)).all() and (info["is_success"] == terminated).all()
    assert torch.allclose(u.get_cube().get_pos(), obs["environment_state"][:, 0:3])
    assert torch.allclose(u._env.cube_2.get_pos(), obs["environment_state"][:, 11:14])
    # same seed -> bit-identical observations (the property tester.py:11-28 probes)
    obs2, _ = env.reset(seed=3)
    assert (obs2["environment_state"] == first).all()
    # the reference's host RandomState stream: cube_1 x draws first, then cube_1 y, cube_2 x, ... (cube_stack.py:123-145)
    rs = np.random.RandomState(3)
    ref = np.zeros((n, 5, 2))
    for k, (xlo, xhi, ylo, yhi) in enumerate(sm.obj_spawn):
        ref[:, k, 0] = rs.uniform(xlo, xhi, size=(n,))
        ref[:, k, 1] = rs.uniform(ylo, yhi, size=(n,))
    es = obs2["environment_state"].cpu().numpy()
    clear = np.ones((n, 5), bool)                             # cubes that touch no other cube in the settle step
    for i in range(5):
        for j in range(5):
            if i != j:
                clear[:, i] &= np.abs(ref[:, i] - ref[:, j]).max(axis=1) > 0.08
    assert clear[:, 0].sum() > n // 2 and clear[:, 1].sum() > n // 2
    assert np.abs(es[clear[:, 0], 0:2] - ref[clear[:, 0], 0]).max() < 1e-6
    assert np.abs(es[clear[:, 1], 11:13] - ref[clear[:, 1], 1]).max() < 1e-6
    env.close()


def test_expert_closed_loop_stacks_on_device(Sim, sm):
    """The collectors' stage machine (hover, descend, grasp, lift, carry, lower, release, retreat) with robot.inverse_kinematics
    and env.step both on the device: a clear majority of well-separated envs ends with cube_1 resting on cube_2."""
    from gym_genesis_b200.tasks import FrankaCubeStackEmptyBatch
    n = 256
    task = FrankaCubeStackEmptyBatch(False, 480, 640, n, (1.0, 1.0), "global", True)
    task.seed(1)
    task.reset()
    dev = task.sim.device
    down = torch.tensor([0.0, 1.0, 0.0, 0.0], device=dev).repeat(n, 1)
    plan = [("c1", 0.2, 0.04, 40), ("c1", 0.105, 0.04, 40), ("c1", 0.105, -0.02, 30), ("lift", 0.25, -0.02, 50),
            ("c2", 0.23, -0.02, 70), ("c2", 0.148, -0.02, 60), ("c2", 0.148, 0.04, 30), ("c2", 0.3, 0.04, 30)]
    c1_0, c2_0 = task.cube_1.get_pos(), task.cube_2.get_pos()
    others = torch.stack([c.get_pos() for c in task.distractor_cubes], 1)
    sep = (c1_0 - c2_0)[:, :2].norm(dim=1) > 0.12
    for o in range(others.shape[1]):
        sep &= ((others[:, o] - c1_0)[:, :2].norm(dim=1) > 0.12) & ((others[:, o] - c2_0)[:, :2].norm(dim=1) > 0.12)
    hold = None
    for name, dz, grip, steps in plan:
        for _ in range(steps):
            c1, c2 = task.cube_1.get_pos(), task.cube_2.get_pos()
            if name == "c1":
                tp = c1 + torch.tensor([0.0, 0.0, dz], device=dev); hold = c1.clone()
            elif name == "lift":
                tp = hold + torch.tensor([0.0, 0.0, dz], device=dev)
            else:
                tp = torch.stack([c2[:, 0], c2[:, 1], torch.full((n,), 0.02 + dz, device=dev)], 1)
            qt = task.robot.inverse_kinematics(link=task.eef, pos=tp, quat=down)
            act = qt.clone()
            act[:, 7:] = grip
            task.step(act)
    rew = task.compute_reward()
    assert sep.sum().item() >= 40
    rate = rew[sep].mean().item()
    assert rate > 0.5, (rate, int(sep.sum()))
    assert torch.isfinite(task.sim.get_state()[0]).all()


# ---- SO-101 CubeStackBatch ----------------------------------------------------------------------
@pytest.fixture(scope="module")
def so():
    from gym_genesis_b200.model_so101 import so101_cube_stack_model
    from oracle.binding import Oracle
    m = so101_cube_stack_model()
    return m, Oracle(m, np.float64)


def test_so101_stack_single_substep(Sim, so):
    from test_stack_cpu import _so101_states
    m, orc = so
    n = 128
    rng = np.random.default_rng(41)
    Q, QD, T = _so101_states(m, rng, n)
    QA = np.zeros_like(QD)
    sim = Sim(m, n)
    assert sim.agent_dim == 6 and sim.state.shape[0] == 3 * 6 + 19 * 5
    q, qd, qa = _one_substep(sim, Q, QD, QA, T)
    qo, qdo, qao = Q.copy(), QD.copy(), QA.copy()
    info, _ = orc.step(qo, qdo, qao, T, iterations=100, tol=1e-12)
    nc, it = sim.step_info()
    assert (nc.cpu().numpy() == info[:, 0]).all()
    eq, ev = stack_substep_errors(Q, QD, qo, qdo, q, qd, _asm(orc, Q, QD, T))
    assert eq.max() < TOL_Q_ABS and ev.max() < 1.0, (eq.max(), ev.max(), int(np.argmax(ev)))
    a, e, r, t = orc.obs(q)
    assert (sim.obs_agent.cpu().numpy() == f32(q[:, :6])).all()                      # agent_pos = joint positions, bit-exact
    assert np.abs(_np(sim.obs_env) - e).max() < 2e-6
    assert (sim.terminated.cpu().numpy().astype(bool) == t).all() and t[2]


def test_so101_stack_gym_surface_and_reset_stream(so):
    import gym_genesis  # noqa: F401
    from gym_genesis_b200.gym_api import gym
    m, orc = so
    n = 40
    env = gym.make("gym_genesis/CubeStack-v0", robot="so101", num_envs=n)
    u = env.unwrapped
    obs, _ = env.reset(seed=11)
    assert obs["agent_pos"].shape == (n, 6) and obs["environment_state"].shape == (n, 14)
    # the reference's host stream (cube_stack_batch.py:70-97): per-env rejection loop, then vector draws per distractor
    rs = np.random.RandomState(11)
    xy = np.zeros((n, 5, 2))
    for b in range(n):
        while True:
            x1, y1, x2, y2 = rs.uniform(-0.3, -0.1), rs.uniform(-0.1, 0.1), rs.uniform(-0.3, -0.1), rs.uniform(-0.1, 0.1)
            if ((x2 - x1) ** 2 + (y2 - y1) ** 2) ** 0.5 >= 0.06:
                xy[b, 0], xy[b, 1] = (x1, y1), (x2, y2)
                break
    for k in range(2, 5):
        xy[:, k, 0] = rs.uniform(-0.35, 0.0, size=(n,)); xy[:, k, 1] = rs.uniform(-0.2, 0.2, size=(n,))
    # oracle: reset + one settle step holding the home pose
    qo, qdo, qao = orc.new_state(n)
    orc.reset(qo, qdo, qao, f32(xy).astype(np.float64))
    orc.step(qo, qdo, qao, np.tile(m.home_q, (n, 1)), iterations=100, tol=1e-12)
    q = _np(u._env.sim.get_state()[0])
    assert np.abs(q - qo).max() < 5e-5, np.abs(q - qo).max()
    a, e, r, t = orc.obs(q)
    assert np.abs(_np(obs["environment_state"]) - e).max() < 2e-6
    act = torch.as_tensor(f32(np.tile(m.home_q, (n, 1))), device="cuda")
    obs, reward, terminated, truncated, info = u.step(act)
    assert reward.shape == (n,) and (terminated == (reward == 1)).all() and not truncated.any()
    assert torch.isfinite(obs["environment_state"]).all()
    env.close()


@pytest.mark.parametrize("robot,adim", [("so101", 6), ("franka", 9)])
def test_stack_unbatched_variants(robot, adim):
    """num_envs = 0 selects the unbatched task (reference env.py:114-120): same physics, env dimension dropped."""
    import gym_genesis  # noqa: F401
    from gym_genesis_b200.gym_api import gym
    env = gym.make("gym_genesis/CubeStack-v0", robot=robot, num_envs=0)
    obs, _ = env.reset(seed=11)
    assert obs["agent_pos"].shape == (adim,) and obs["environment_state"].shape == (14,)
    ref = gym.make("gym_genesis/CubeStack-v0", robot=robot, num_envs=1)
    obs1, _ = ref.reset(seed=11)
    assert torch.equal(obs1["environment_state"][0], obs["environment_state"])      # identical reset stream at B = 1
    a = torch.as_tensor(f32(env.unwrapped._env.model.home_q), device="cuda")
    o2, reward, terminated, truncated, info = env.unwrapped.step(a)
    assert o2["agent_pos"].shape == (adim,) and reward.dim() == 0
    env.close(); ref.close()


def test_sharding_invariance(Sim, sm):
    """Rank r of G owns global envs [r n, (r + 1) n) (`first_env_id`): spawn draws and Bernoulli resets are keyed by the GLOBAL env
    id, so two shards of n envs reproduce one simulator of 2 n envs bit for bit (no collective on the step path)."""
    n = 96
    kw = dict(auto_reset=True, reset_prob=0.05, seed=77)
    whole = Sim(sm, 2 * n, **kw)
    parts = [Sim(sm, n, first_env_id=r * n, **kw) for r in range(2)]
    g = torch.Generator(device="cuda").manual_seed(1)
    home = torch.as_tensor(f32(FRANKA_HOME), device="cuda")
    for s in [whole] + parts:
        s.reset(None, None)
    for _ in range(12):
        a = home + (torch.rand((2 * n, 9), device="cuda", generator=g) - 0.5) * 0.5
        whole.step(a)
        for r in range(2):
            parts[r].step(a[r * n:(r + 1) * n].contiguous())
    qw, qdw, _ = whole.get_state()
    for r in range(2):
        q, qd, _ = parts[r].get_state()
        assert torch.equal(q, qw[r * n:(r + 1) * n]) and torch.equal(qd, qdw[r * n:(r + 1) * n])
        assert torch.equal(parts[r].episode, whole.episode[r * n:(r + 1) * n])
        assert torch.equal(parts[r].obs_env, whole.obs_env[r * n:(r + 1) * n])
    assert whole.episode.max().item() > 1                     # Bernoulli resets did happen


def test_so101_stack_rollout(Sim, so):
    m, orc = so
    n = 16
    rng = np.random.default_rng(9)
    xy = np.zeros((n, 5, 2), np.float32)
    for k, (xlo, xhi, ylo, yhi) in enumerate(m.obj_spawn):
        xy[:, k, 0] = rng.uniform(xlo, xhi, n); xy[:, k, 1] = rng.uniform(ylo, yhi, n)
    sim = Sim(m, n)
    sim.reset(None, xy)
    q, qd, qa = (_np(t) for t in sim.get_state())
    tgt = np.tile(m.home_q, (n, 1))
    for _ in range(60):
        tgt = tgt + rng.normal(0, 0.01, tgt.shape)
        t32 = f32(tgt)
        sim.step(torch.as_tensor(t32))
        orc.step(q, qd, qa, t32.astype(np.float64), iterations=100, tol=1e-10)
    qg = _np(sim.get_state()[0])
    pos_cols = list(range(6)) + [6 + 7 * k + c for k in range(5) for c in range(3)]
    err = np.abs(qg[:, pos_cols] - q[:, pos_cols]).max()
    assert err < ROLLOUT_POS_TOL, err


@pytest.mark.parametrize("robot", ["franka", "so101"])
def test_step_host_zero_copy_is_bit_identical_to_device_step(Sim, robot):
    """gg_stack_step_host (pinned host action read in place by the prep kernel, reward / flags mirrored into pinned host memory by
    the finish kernel) against gg_stack_step with a device-resident action: bit-identical state, observations and flags."""
    from gym_genesis_b200.model import get_model
    m = get_model(robot, "cube_stack")
    n = 1000
    rng = np.random.default_rng(78)
    a = Sim(m, n, seed=3, auto_reset=True, reset_prob=0.02, time_limit=True)
    b = Sim(m, n, seed=3, auto_reset=True, reset_prob=0.02, time_limit=True)
    a.reset(None, None); b.reset(None, None)
    for k in range(10):
        act = f32(np.tile(m.home_q, (n, 1)) + rng.uniform(-0.2, 0.2, (n, m.nu)))
        a.step(torch.as_tensor(act, device=a.device))
        pinned = torch.as_tensor(act).pin_memory() if k % 2 == 0 else act          # pinned: zero-copy; numpy: staged copy
        rew, term, trunc = b.step_host(pinned)
        assert torch.equal(rew, a.reward.cpu()) and torch.equal(term, a.terminated.cpu()) and torch.equal(trunc, a.truncated.cpu())
        assert torch.equal(a.obs_agent, b.obs_agent) and torch.equal(a.obs_env, b.obs_env)
    for x, y in zip(a.get_state(), b.get_state()):
        assert torch.equal(x, y)
    import ctypes as C
    from gym_genesis_b200 import cabi
    from gym_genesis_b200.sim import GgError
    bad = torch.zeros((n, m.nu), dtype=torch.float32)                               # pageable host memory must be rejected
    io = cabi.GgHostIO(bad.data_ptr(), 0, 0, 0, 1)
    with pytest.raises(GgError, match="pinned"):
        b._check(b.lib.gg_stack_step_host(b._h, C.byref(io), b.obs_agent.data_ptr(), b.obs_env.data_ptr(), None, None, None, None, b._stream()))


def test_gym_surface_host_io_cube_stack():
    import gym_genesis  # noqa: F401
    from gym_genesis_b200.gym_api import gym
    n = 64
    env = gym.make("gym_genesis/CubeStack-v0", robot="franka", num_envs=n, host_io=True, scene="empty").unwrapped
    obs, info = env.reset(seed=0)
    act = np.tile(FRANKA_HOME.astype(np.float32), (n, 1))
    obs, reward, terminated, truncated, info = env.step(act)                           # numpy action, as in the reference README
    assert reward.device.type == "cpu" and reward.shape == (n,) and terminated.dtype == torch.bool and not truncated.any()
    assert obs["agent_pos"].is_cuda and obs["environment_state"].shape == (n, 14)
    assert torch.equal(reward, env._env.sim.reward.cpu())
    env.close()


# ---- FrankaCubeStackKitchenBatch: what the reference wires to CubeStack-v0 + franka (scaled Panda on the kitchen island) ----------
def test_kitchen_variant_substep_and_surface(Sim):
    from gym_genesis_b200.model import franka_cube_stack_kitchen_model
    from oracle.binding import Oracle
    m = franka_cube_stack_kitchen_model()
    orc = Oracle(m, np.float64)
    n = 96
    rng = np.random.default_rng(61)
    sim = Sim(m, n)
    Q, QD, T = stack_clutter_states(m, sim.gg_model, rng, n)
    Q2, QD2, T2 = stack_random_states(m, rng, n)
    Q2[:, :9] = np.clip(Q2[:, :9], np.array(list(sim.gg_model.lim_lo)), np.array(list(sim.gg_model.lim_hi)))
    for (q0, qd0, t0) in ((Q, QD, T), (f32(Q2).astype(np.float64), QD2, T2)):
        qa0 = np.zeros_like(qd0)
        q, qd, qa = _one_substep(sim, q0, qd0, qa0, t0)
        qo, qdo, qao = q0.copy(), qd0.copy(), qa0.copy()
        info, _ = orc.step(qo, qdo, qao, t0, iterations=300, tol=1e-12)
        assert (sim.step_info()[0].cpu().numpy() == info[:, 0]).all()
        eq, ev = stack_substep_errors(q0, qd0, qo, qdo, q, qd, _asm(orc, q0, qd0, t0))
        assert eq.max() < TOL_Q_ABS and ev.max() < 1.0, (eq.max(), ev.max(), int(np.argmax(ev)))
    import gym_genesis  # noqa: F401
    from gym_genesis_b200.gym_api import gym
    env = gym.make("gym_genesis/CubeStack-v0", robot="franka", num_envs=32)            # the reference's mapping: kitchen variant
    obs, _ = env.reset(seed=4)
    es = obs["environment_state"].cpu().numpy()
    assert obs["agent_pos"].shape == (32, 9) and es.shape == (32, 14)
    assert (es[:, 0] > -0.31).all() and (es[:, 0] < -0.09).all() and np.abs(es[:, 2] - (m.plane_z + 0.02)).max() < 2e-3   # :71, :67
    assert np.abs(obs["agent_pos"][:, 2].cpu().numpy() - 1.037).max() < 5e-3                                            # scaled arm at home
    env.close()


def test_full_size_oracle_parity_with_masked_resets_every_step(Sim, sm, orc):
    """CubeStack at 16384 envs (the pick path's BASELINE sizes applied to row 8f-1): Bernoulli(1/200) partial resets every step +
    auto-reset + time limit.  Every step: reset decision, counters, reward / terminated / truncated and observations of ALL envs
    reproduced on the host (flags bit-exact from the GPU's own post-step state through the oracle's float32 reward rule); a subset of
    256 envs -- the last block, every env that reset, random others -- advanced by the fp64 oracle from the GPU's pre-step state."""
    from helpers import philox4x32_10_np, u01_np
    n, steps, seed, first = 16384, 8, (7 << 32) | 99, 4096
    rp = 1.0 / 200
    rng = np.random.default_rng(5)
    sim = Sim(sm, n, seed=seed, reset_prob=rp, auto_reset=True, time_limit=True, first_env_id=first)
    sim.reset(None, None)
    nu, K = sm.nu, sm.n_obj
    gid = (first + np.arange(n)).astype(np.uint32)
    done_prev = np.zeros(n, bool)
    home32 = f32(sm.home_q)
    n_reset = n_checked = n_soft = 0
    from oracle.binding import Oracle
    orc32 = Oracle(sm, np.float32)
    for k in range(steps):
        q0, qd0, qa0 = (_np(t) for t in sim.get_state())
        sc0, ep0 = sim.step_count.cpu().numpy().copy(), sim.episode.cpu().numpy().copy()
        act = home32 + rng.uniform(-0.3, 0.3, (n, nu)).astype(np.float32)
        act[:, 7:] = rng.uniform(-0.02, 0.04, (n, 2)).astype(np.float32)
        sim.step(torch.as_tensor(act))
        o = philox4x32_10_np(gid, np.full(n, k, np.uint32), np.full(n, 0x52535421, np.uint32), np.zeros(n, np.uint32),
                             np.uint32(seed & 0xFFFFFFFF), np.uint32(seed >> 32))
        do_reset = done_prev | (u01_np(o[0]) < np.float32(rp))
        q1, qd1, _ = (_np(t) for t in sim.get_state())
        sc1, ep1 = sim.step_count.cpu().numpy(), sim.episode.cpu().numpy()
        assert np.array_equal(ep1, ep0 + do_reset) and np.array_equal(sc1, np.where(do_reset, 0, sc0 + 1)), f"step {k}: counters"
        a, e, r, t = orc.obs(q1)
        assert np.array_equal(sim.terminated.cpu().numpy().astype(bool), t) and np.array_equal(sim.reward.cpu().numpy(), r.astype(np.float32))
        trunc = sc1 >= sm.horizon
        assert np.array_equal(sim.truncated.cpu().numpy().astype(bool), trunc)
        assert np.abs(_np(sim.obs_agent) - a).max() < 2e-6 and np.abs(_np(sim.obs_env) - e).max() < 2e-6
        done_prev = t | trunc
        ridx = np.nonzero(do_reset)[0]
        n_reset += len(ridx)
        idx = np.unique(np.concatenate([np.arange(n - 32, n), ridx[:64], rng.choice(n, 200, replace=False)]))
        is_r = do_reset[idx]
        Q, QD, QA = q0[idx].copy(), qd0[idx].copy(), qa0[idx].copy()
        T = act[idx].astype(np.float64)
        for j in np.nonzero(is_r)[0]:
            Q[j, :nu] = sm.home_q
            for c in range(K):
                xy = orc.spawn_obj(seed, int(gid[idx[j]]), int(ep0[idx[j]]), c)
                Q[j, nu + 7 * c: nu + 7 * c + 7] = [xy[0], xy[1], np.float32(sm.spawn_z), *f32(sm.spawn_quat)]
            QD[j] = 0; QA[j] = 0
            T[j] = home32
        Q, QD = f32(Q).astype(np.float64), f32(QD).astype(np.float64)
        qo, qdo, qao = Q.copy(), QD.copy(), QA.copy()
        orc.step(qo, qdo, qao, T, iterations=100, tol=1e-12)
        asm = _asm(orc, Q, QD, T)
        eq, ev = stack_substep_errors(Q, QD, qo, qdo, q1[idx], qd1[idx], asm)
        # At this scale a few envs per step hold configurations that float32 itself cannot resolve to the stated bound (two axis-aligned
        # cubes touching face to face with 0.2 mm of overlap: the torque about z is a difference of nearly equal lever arms).  The
        # yardstick for those is the oracle's OWN float32 build on the same state: the kernel may deviate from the fp64 oracle by at
        # most twice what plain float32 arithmetic on the oracle's formulation does.
        over = np.nonzero((ev >= 1.0) | (eq >= TOL_Q_ABS))[0]
        n_soft += len(over)
        if len(over):
            q32, qd32, qa32 = f32(Q[over]).copy(), f32(QD[over]).copy(), f32(QA[over]).copy()
            orc32.step(q32, qd32, qa32, f32(T[over]), iterations=100, tol=1e-7)
            eq32, ev32 = stack_substep_errors(Q[over], QD[over], qo[over], qdo[over], q32.astype(np.float64), qd32.astype(np.float64), asm[over])
            assert (ev[over] < 2.0 * np.maximum(ev32, 0.5)).all() and (eq[over] < 2.0 * np.maximum(eq32, 0.5 * TOL_Q_ABS)).all(), \
                (k, ev[over], ev32, eq[over], eq32, idx[over])
        n_checked += len(idx)
    assert n_reset > steps * 50 and n_checked > steps * 250
    assert n_soft <= 0.01 * n_checked                   # the float32-limited configurations are rare (< 1 % of the checked envs)
    assert sim.stats().cpu().numpy()[5] == 0
