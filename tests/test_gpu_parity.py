"""GPU tier: the CUDA path, called through the C ABI (BatchedSim -> libgg_b200.so), against the oracle."""
import numpy as np
import pytest
import torch

from gym_genesis_b200 import model as mdl
from helpers import (ROLLOUT_POS_TOL, TOL_Q_ABS, f32, pick_clutter_states, random_states, scripted_grasp_states, substep_errors)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def Sim():
    from gym_genesis_b200.sim import BatchedSim
    return BatchedSim


def _np(t):
    return t.detach().cpu().numpy().astype(np.float64)


def _a_smooth_scale(o, Q, QD, T):
    return np.array([np.abs(o.dynamics(Q[k], QD[k], T[k])["a_smooth"]).max() * 0.01 for k in range(Q.shape[0])])


def _one_substep(sim, Q, QD, QA, T):
    sim.set_state(Q, QD, QA)
    sim.step(torch.as_tensor(f32(T)))
    q, qd, qa = sim.get_state()
    return _np(q), _np(qd), _np(qa)


@pytest.mark.parametrize("on_ground", [True, False])
def test_single_substep_random_states(Sim, franka_model, oracle64, on_ground):
    n = 256
    rng = np.random.default_rng(10 + on_ground)
    Q, QD, T = random_states(franka_model, rng, n, cube_on_ground=on_ground)
    QA = np.zeros_like(QD)
    sim = Sim(franka_model, n)
    q, qd, qa = _one_substep(sim, Q, QD, QA, T)
    qo, qdo, qao = Q.copy(), QD.copy(), QA.copy()
    oracle64.step(qo, qdo, qao, T, iterations=100, tol=1e-12)
    eq, ev = substep_errors(Q, QD, qo, qdo, q, qd, _a_smooth_scale(oracle64, Q, QD, T))
    assert eq.max() < TOL_Q_ABS and ev.max() < 1.0, (eq.max(), ev.max())
    # observations / reward / flags of the step against the oracle on the GPU's own state (bit-exact flags)
    a, e, r, t = oracle64.obs(q)
    assert np.abs(_np(sim.obs_agent) - a).max() < 2e-6 and np.abs(_np(sim.obs_env) - e).max() < 2e-6
    assert (sim.terminated.cpu().numpy().astype(bool) == t).all()
    assert (sim.reward.cpu().numpy() == r.astype(np.float32)).all()
    assert not sim.truncated.any()


def test_single_substep_scripted_grasp_states(Sim, franka_model, oracle64):
    """Contact-rich states (pads on the ground, pads on the cube, up to the 64-contact cap)."""
    states, reward = scripted_grasp_states(franka_model, oracle64, stride=2)
    assert reward == 1.0
    Q, QD, QA, T = (np.concatenate([f32(s[k]) for s in states]).astype(np.float64) for k in range(4))
    sim = Sim(franka_model, Q.shape[0])
    q, qd, qa = _one_substep(sim, Q, QD, QA, T)
    qo, qdo, qao = Q.copy(), QD.copy(), QA.copy()
    info, _ = oracle64.step(qo, qdo, qao, T, iterations=100, tol=1e-12)
    assert info[:, 0].max() >= 32
    eq, ev = substep_errors(Q, QD, qo, qdo, q, qd, _a_smooth_scale(oracle64, Q, QD, T))
    assert eq.max() < TOL_Q_ABS and ev.max() < 1.0, (eq.max(), ev.max())


def test_single_substep_gripper_cube_clutter(Sim, franka_model, oracle64):
    """1024 states with the cube at random poses around the fingers (overlapping pads and / or plane): the warp-per-env solver teams
    next to the register-resident path inside the same blocks."""
    sim = Sim(franka_model, 1024)
    Q, QD, T = pick_clutter_states(franka_model, sim.gg_model, np.random.default_rng(56), 1024)
    QA = np.zeros_like(QD)
    q, qd, qa = _one_substep(sim, Q, QD, QA, T)
    qo, qdo, qao = Q.copy(), QD.copy(), QA.copy()
    info, _ = oracle64.step(qo, qdo, qao, T, iterations=300, tol=1e-12)
    assert (info[:, 0] > 4).sum() > 200
    eq, ev = substep_errors(Q, QD, qo, qdo, q, qd, _a_smooth_scale(oracle64, Q, QD, T))
    assert eq.max() < TOL_Q_ABS and ev.max() < 1.0, (eq.max(), ev.max(), int(np.argmax(ev)))


def test_rollout_100_steps(Sim, franka_model, oracle64):
    """100 control steps, smooth random-walk targets, cube resting: positions stay within the stated bound."""
    n = 32
    rng = np.random.default_rng(20)
    sim = Sim(franka_model, n)
    xy = np.stack([rng.uniform(0.45, 0.8, n), rng.uniform(-0.25, 0.25, n)], 1).astype(np.float32)
    sim.reset(None, xy)
    q, qd, qa = (_np(t) for t in sim.get_state())
    tgt = np.tile(mdl.FRANKA_HOME, (n, 1))
    worst = 0.0
    for k in range(100):
        tgt = tgt + rng.normal(0, 0.02, tgt.shape)
        tgt[:, 7:] = np.clip(tgt[:, 7:], -0.02, 0.04)
        t32 = f32(tgt)
        sim.step(torch.as_tensor(t32))
        oracle64.step(q, qd, qa, t32.astype(np.float64), iterations=100, tol=1e-12)
    qg = _np(sim.get_state()[0])
    worst = np.abs(qg[:, :12] - q[:, :12]).max()
    assert worst < ROLLOUT_POS_TOL, worst
    assert worst < 1e-4          # measured far below the stated bound in free space / resting contact


def test_reset_reproduces_reference_sequence(Sim, franka_model, oracle64):
    """Full reset = spawn at the host RandomState draws, home pose, hold-home target, ONE settle step, obs
    (franka/cube_pick.py:86-112); the post-reset cube quaternion is the literal (0,0,0,1) of the reference."""
    n = 128
    sim = Sim(franka_model, n)
    r = np.random.RandomState(7)
    xy = np.stack([r.uniform(0.45, 0.80, n), r.uniform(-0.25, 0.25, n)], 1).astype(np.float32)
    sim.reset(None, xy)
    Q, QD, QA = oracle64.new_state(n)
    oracle64.reset(Q, QD, QA, xy.astype(np.float64))
    oracle64.step(Q, QD, QA, np.tile(mdl.FRANKA_HOME, (n, 1)))
    q, qd, _ = (_np(t) for t in sim.get_state())
    assert np.abs(q - Q).max() < TOL_Q_ABS
    env = _np(sim.obs_env)
    assert np.array_equal(env[:, 0:2].astype(np.float32), xy)            # spawn xy untouched by the settle step
    assert np.abs(env[:, 3:7] - [0, 0, 0, 1]).max() < 1e-6
    assert (sim.step_count.cpu().numpy() == 0).all() and (sim.episode.cpu().numpy() == 1).all()


def test_masked_partial_reset_is_bit_exact_elsewhere(Sim, franka_model, oracle64):
    n = 512
    rng = np.random.default_rng(30)
    sim_a, sim_b = Sim(franka_model, n, seed=123), Sim(franka_model, n, seed=123)
    xy = np.stack([rng.uniform(0.45, 0.8, n), rng.uniform(-0.25, 0.25, n)], 1).astype(np.float32)
    act = torch.as_tensor(f32(mdl.FRANKA_HOME + rng.uniform(-0.2, 0.2, (n, 9))), device="cuda")
    mask = torch.as_tensor(rng.random(n) < 0.25, device="cuda")
    for s in (sim_a, sim_b):
        s.reset(None, xy)
        s.step(act)
    sim_a.step(act)                       # nobody resets
    sim_b.step(act, reset_mask=mask)      # 25 % reset instead of stepping
    qa_, qb_ = sim_a.get_state()[0], sim_b.get_state()[0]
    keep = ~mask
    assert torch.equal(qa_[keep], qb_[keep])                                  # untouched envs: bit-identical
    assert torch.equal(sim_a.obs_env[keep], sim_b.obs_env[keep])
    # reset envs: spawn xy from the Philox stream (seed, global env id, episode), bit-exact against the oracle
    idx = torch.nonzero(mask).flatten().cpu().numpy()
    env = sim_b.obs_env.cpu().numpy()
    for e in idx[:64]:
        ref = oracle64.spawn_xy(123, int(e), 1)
        assert env[e, 0:2].tobytes() == ref.tobytes()
    assert (sim_b.step_count.cpu().numpy()[idx] == 0).all() and (sim_b.episode.cpu().numpy()[idx] == 2).all()
    assert (sim_b.step_count.cpu().numpy()[keep.cpu().numpy()] == 2).all()


def test_flags_success_and_auto_reset(Sim, franka_model):
    n = 64
    sim = Sim(franka_model, n, auto_reset=True, time_limit=True)
    sim.reset(None, np.tile(np.array([[0.6, 0.0]], np.float32), (n, 1)))
    q, qd, qa = sim.get_state()
    q[:16, 11] = 0.5                      # lift 16 cubes above the success height
    sim.set_state(q, qd, qa)
    home = torch.as_tensor(np.tile(f32(mdl.FRANKA_HOME), (n, 1)), device="cuda")
    sim.step(home)
    term = sim.terminated.cpu().numpy().astype(bool)
    assert term[:16].all() and not term[16:].any()
    assert (sim.reward.cpu().numpy() == term.astype(np.float32)).all()
    sim.step(home)                        # finished envs reset on the next step (episode counter advances)
    assert (sim.episode.cpu().numpy()[:16] == 2).all() and (sim.episode.cpu().numpy()[16:] == 1).all()
    assert not sim.terminated.cpu().numpy()[:16].any()
    st = sim.stats().cpu().numpy()
    assert st[1] == 16 and st[2] == 16 and st[5] == 0
    # time limit: truncated exactly at horizon
    sim2 = Sim(franka_model, 4, time_limit=True)
    sim2.reset(None, np.tile(np.array([[0.6, 0.0]], np.float32), (4, 1)))
    h4 = home[:4].contiguous()
    for k in range(200):
        sim2.step(h4)
        assert bool(sim2.truncated.all()) == (k == 199)


def test_same_seed_determinism(Sim, franka_model):
    """The one property the reference author probed (tester.py:11-28): same seed -> same observations (here: bit-identical)."""
    n = 256
    outs = []
    for _ in range(2):
        sim = Sim(franka_model, n, seed=42, reset_prob=0.02)
        sim.reset(None, None)
        g = torch.Generator(device="cpu").manual_seed(0)
        for k in range(20):
            act = torch.as_tensor(mdl.FRANKA_HOME, dtype=torch.float32) + 0.3 * (torch.rand((n, 9), generator=g) - 0.5)
            sim.step(act)
        outs.append((sim.obs_agent.clone(), sim.obs_env.clone(), sim.episode.clone()))
    assert all(torch.equal(a, b) for a, b in zip(*outs))
    assert outs[0][2].max() > 1           # random resets happened


def test_full_size_invariants(Sim, franka_model):
    """BASELINE config sizes (16384 envs, randomized poses, masked resets every step): size-independent properties."""
    n = 16384
    sim = Sim(franka_model, n, seed=1, reset_prob=1.0 / 200, auto_reset=True)
    sim.reset(None, None)
    g = torch.Generator(device="cuda").manual_seed(0)
    home = torch.as_tensor(mdl.FRANKA_HOME, dtype=torch.float32, device="cuda")
    for k in range(50):
        act = home + 0.6 * (torch.rand((n, 9), generator=g, device="cuda") - 0.5)
        act[:, 7:] = -0.02 + 0.06 * torch.rand((n, 2), generator=g, device="cuda")
        sim.step(act)
    q, qd, qa = sim.get_state()
    assert torch.isfinite(q).all() and torch.isfinite(qd).all()
    assert (q[:, 12:16].norm(dim=1) - 1).abs().max() < 1e-5               # unit cube quaternion
    assert q[:, 11].min() > 0.0195 and q[:, 11].max() < 0.021             # cubes rest on the plane
    lo = torch.as_tensor(franka_model.dof_table()["lim_lo"][:9], dtype=torch.float32, device="cuda") - 0.01
    hi = torch.as_tensor(franka_model.dof_table()["lim_hi"][:9], dtype=torch.float32, device="cuda") + 0.01
    assert ((q[:, :9] > lo) & (q[:, :9] < hi)).all()                      # soft joint limits hold
    env = sim.obs_env
    assert torch.allclose(env[:, 10], env[:, 7:10].norm(dim=1), atol=1e-6)
    st = sim.stats().cpu().numpy()
    assert st[5] == 0 and st[0] > 0.98 * n * 50 * (1 - 1 / 200.0)


def test_gym_surface_end_to_end():
    import gym_genesis  # noqa: F401
    from gym_genesis_b200.gym_api import gym
    env = gym.make("gym_genesis/CubePick-v0", robot="franka", num_envs=8)
    obs, info = env.reset(seed=3)
    assert info == {"is_success": [False] * 8}
    assert obs["agent_pos"].shape == (8, 9) and obs["environment_state"].shape == (8, 11)
    assert obs["agent_pos"].dtype == torch.float32 and obs["agent_pos"].is_cuda
    r = np.random.RandomState(3)
    x = r.uniform(0.45, 0.80, size=(8,)); y = r.uniform(-0.25, 0.25, size=(8,))
    assert np.array_equal(obs["environment_state"][:, :2].cpu().numpy(), np.stack([x, y], 1).astype(np.float32))
    base = env.unwrapped
    act = np.tile(f32(mdl.FRANKA_HOME), (8, 1))
    obs, reward, terminated, truncated, info = base.step(act)
    assert reward.shape == (8,) and reward.dtype == torch.float32 and terminated.dtype == torch.bool and truncated.dtype == torch.bool
    assert not bool(terminated.any()) and not bool(truncated.any()) and info["is_success"].shape == (8,)
    base.push()
    assert base.get_robot().get_dofs_position().shape == (8, 9) and base.get_cube().get_pos().shape == (8, 3)
    with pytest.raises(ValueError):
        base.step(np.zeros((8, 7), np.float32))
    with pytest.raises(NotImplementedError):
        gym.make("gym_genesis/CubePick-v0", robot="franka", num_envs=2, enable_pixels=True)
    env.close()


def test_so101_substep_and_full_size(Sim, so101_model, so101_oracle64):
    """BASELINE config 4: SO-101 articulation (6 hinge dofs), 16384 envs."""
    m, o = so101_model, so101_oracle64
    rng = np.random.default_rng(40)
    n = 256
    Q, QD, T = random_states(m, rng, n, cube_on_ground=True)
    QA = np.zeros_like(QD)
    sim = Sim(m, n)
    assert sim.agent_dim == 8 and sim.nu == 6
    q, qd, qa = _one_substep(sim, Q, QD, QA, T)
    qo, qdo, qao = Q.copy(), QD.copy(), QA.copy()
    o.step(qo, qdo, qao, T, iterations=100, tol=1e-12)
    sc = np.array([np.abs(o.dynamics(Q[k], QD[k], T[k])["a_smooth"]).max() * 0.01 for k in range(n)])
    eq, ev = substep_errors(Q, QD, qo, qdo, q, qd, sc)
    assert eq.max() < TOL_Q_ABS and ev.max() < 1.0, (eq.max(), ev.max())
    a, e, r, t = o.obs(q)
    assert np.abs(_np(sim.obs_agent) - a).max() < 2e-6 and (sim.terminated.cpu().numpy().astype(bool) == t).all()
    # full size: reset (no settle step for the SO-101 pick task, so101/cube_pick.py:81) then 50 steps with masked resets
    n = 16384
    sim = Sim(m, n, seed=2, reset_prob=1.0 / 200, auto_reset=True)
    sim.reset(None, None)
    env0 = sim.obs_env.clone()
    assert (env0[:, 0] >= -0.32).all() and (env0[:, 0] <= -0.28).all() and (env0[:, 1].abs() <= 0.05).all()
    assert torch.allclose(env0[:, 3:7], torch.tensor([1.0, 0, 0, 0], device="cuda").expand(n, 4))
    g = torch.Generator(device="cuda").manual_seed(0)
    home = torch.as_tensor(m.home_q, dtype=torch.float32, device="cuda")
    for k in range(50):
        sim.step(home + 0.2 * (torch.rand((n, 6), generator=g, device="cuda") - 0.5))
    q, qd, _ = sim.get_state()
    assert torch.isfinite(q).all() and torch.isfinite(qd).all()
    assert (q[:, 9:13].norm(dim=1) - 1).abs().max() < 1e-5
    # Resting depth follows the closed form m g = 16 D K imp depth.  With mu = 5 (so101/cube_pick.py:38-39) the pyramidal
    # regulariser R = 2 mu^2 (1+mu^2) (1-imp)/imp w makes the normal direction 325x softer than at mu = 1, so the cube
    # sinks ~16 mm (a property of the restated constraint model, flagged in DESIGN.md section 2).
    mu, w, K, imp = 5.0, 1 / 0.0128, 1 / (0.95 ** 2 * 0.02 ** 2), 0.95
    depth = 0.0128 * 9.81 / (16 * K * imp / (2 * mu * mu * (1 - imp) / imp * w * (1 + mu * mu)))
    z = q[:, 8].double()
    settled = sim.step_count > 30
    assert settled.float().mean() > 0.8
    assert ((z[settled] - (m.plane_z + 0.02 - depth)).abs() < 0.1 * depth).float().mean() > 0.99
    assert z.min() > m.plane_z - 0.001 and z.max() < m.plane_z + 0.022
    assert sim.stats().cpu().numpy()[5] == 0


def test_scripted_grasp_rollout_open_loop(Sim, franka_model):
    """Same as the CPU-tier test, on the GPU at scale: 32 recorded expert rollouts tiled over 4096 envs."""
    import os
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "grasp_traj.npz"))
    K, rep = d["xy"].shape[0], 128
    n = K * rep
    sim = Sim(franka_model, n)
    sim.reset(None, np.tile(d["xy"], (rep, 1)))
    acts = torch.as_tensor(np.tile(d["actions"], (rep, 1, 1)), device="cuda")
    succ = torch.zeros(n, dtype=torch.bool, device="cuda")
    for t in range(200):
        sim.step(acts[:, t].contiguous())
        succ |= sim.terminated
        if t in (79, 99, 199):
            q = sim.get_state()[0].cpu().numpy()
            er = np.abs(q[:, :9] - np.tile(d["qrob"][:, t], (rep, 1))).max(1)
            ec = np.abs(q[:, 9:12] - np.tile(d["cube"][:, t], (rep, 1))).max(1)
            if t == 79:
                assert er.max() < 5e-6 and ec.max() < 1e-6, (er.max(), ec.max())
            if t == 99:
                assert er.max() < ROLLOUT_POS_TOL and ec.max() < ROLLOUT_POS_TOL and np.median(ec) < 1e-5, (er.max(), ec.max())
            if t == 199:
                assert (ec < 5e-3).mean() >= 0.9
    ref = np.tile(d["reward"].max(1) > 0, rep)
    s = succ.cpu().numpy()
    assert (s == ref).mean() >= 0.9
    assert (s.reshape(rep, K) == s[:K]).all()           # identical inputs in different warps give identical results
    st = sim.stats().cpu().numpy()
    assert st[5] == 0 and st[7] / st[0] > 10            # no non-finite envs; contact-rich on average


def test_nonfinite_env_is_respawned_and_neighbours_untouched(Sim, franka_model, oracle64):
    """A diverged env (non-finite state after the substep) is flagged, re-spawned at the home pose and counted;
    every other env of the same warp/block steps exactly as if nothing had happened."""
    n = 96
    rng = np.random.default_rng(77)
    Q, QD, T = random_states(franka_model, rng, n, cube_on_ground=True)
    QA = np.zeros_like(QD)
    ref = Sim(franka_model, n)
    q_ref, qd_ref, _ = _one_substep(ref, Q, QD, QA, T)
    QDb = QD.copy()
    QDb[5, 2] = np.nan             # robot joint velocity of env 5
    QDb[40, 10] = np.inf           # cube linear velocity of env 40
    sim = Sim(franka_model, n)
    q, qd, _ = _one_substep(sim, Q, QDb, QA, T)
    ok = np.ones(n, bool); ok[[5, 40]] = False
    assert (q[ok] == q_ref[ok]).all() and (qd[ok] == qd_ref[ok]).all()               # bit-exact elsewhere
    for k in (5, 40):
        assert np.isfinite(q[k]).all() and np.isfinite(qd[k]).all()
        assert np.abs(q[k, :9] - franka_model.home_q).max() < 1e-6 and np.abs(qd[k]).max() == 0.0
        assert abs(q[k, 11] - 0.02) < 1e-6
    a, e, r, t = oracle64.obs(q)
    assert np.abs(_np(sim.obs_agent) - a).max() < 2e-6 and np.abs(_np(sim.obs_env) - e).max() < 2e-6
    st = sim.stats().cpu().numpy()
    assert st[5] == 2 and st[0] == n - 2
    assert (sim.episode.cpu().numpy()[[5, 40]] == 1).all() and sim.episode.cpu().numpy()[ok].max() == 0


@pytest.mark.parametrize("n", [1000, 1024])
def test_step_host_zero_copy_is_bit_identical_to_device_step(Sim, franka_model, n):
    """gg_step_host (pinned host action read in place by the kernel, reward / flags mirrored into pinned host memory) against
    gg_step with a device-resident action: bit-identical state, observations and flags; n = 1000 exercises the partial tail
    block (scalar action loads) next to full blocks (cp.async staged slab)."""
    rng = np.random.default_rng(77)
    a = Sim(franka_model, n, seed=3, auto_reset=True, reset_prob=0.02, time_limit=True)
    b = Sim(franka_model, n, seed=3, auto_reset=True, reset_prob=0.02, time_limit=True)
    a.reset(None, None); b.reset(None, None)
    for k in range(12):
        act = f32(np.tile(mdl.FRANKA_HOME, (n, 1)) + rng.uniform(-0.3, 0.3, (n, 9)))
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
    bad = torch.zeros((n, 9), dtype=torch.float32)                                   # pageable host memory must be rejected
    io = cabi.GgHostIO(bad.data_ptr(), 0, 0, 0, 1)
    with pytest.raises(GgError, match="pinned"):
        b._check(b.lib.gg_step_host(b._h, C.byref(io), b.obs_agent.data_ptr(), b.obs_env.data_ptr(), None, None, None, None, b._stream()))


def test_gym_surface_host_io():
    import gym_genesis  # noqa: F401
    from gym_genesis_b200.gym_api import gym
    n = 64
    env = gym.make("gym_genesis/CubePick-v0", robot="franka", num_envs=n, host_io=True).unwrapped
    obs, info = env.reset(seed=0)
    act = np.tile(mdl.FRANKA_HOME.astype(np.float32), (n, 1))
    obs, reward, terminated, truncated, info = env.step(act)                           # numpy action, as in the reference README
    assert reward.device.type == "cpu" and reward.shape == (n,) and reward.dtype == torch.float32
    assert terminated.dtype == torch.bool and not terminated.any() and not truncated.any()
    assert obs["agent_pos"].is_cuda and obs["environment_state"].shape == (n, 11)
    assert torch.equal(reward, env._env.sim.reward.cpu())


def test_step_outputs_are_fresh_by_default_and_alias_on_request():
    """the reference's getters return fresh tensors (franka/cube_pick.py:137-157) and its collectors append them step after
    step (examples/franka/pick_cube_state.py:86-100).  Default: what step t returned is untouched by step t + 1; alias_outputs=True is
    the zero-allocation opt-in."""
    import gym_genesis  # noqa: F401
    from gym_genesis_b200.gym_api import gym
    n = 64
    rng = np.random.default_rng(0)
    acts = [f32(mdl.FRANKA_HOME + rng.uniform(-0.3, 0.3, (n, 9))) for _ in range(3)]
    for host_io in (False, True):
        env = gym.make("gym_genesis/CubePick-v0", robot="franka", num_envs=n, host_io=host_io).unwrapped
        obs0, _ = env.reset(seed=1)
        keep0 = {k: v.clone() for k, v in obs0.items()}
        outs = []
        for a in acts:
            o, r, t, tr, info = env.step(torch.as_tensor(a).pin_memory() if host_io else a)
            outs.append((o, r, t, {k: v.clone() for k, v in o.items()}, r.clone()))
        assert all(torch.equal(obs0[k], keep0[k]) for k in obs0)                       # reset's observation survived three steps
        for o, r, t, oc, rc in outs:
            assert all(torch.equal(o[k], oc[k]) for k in o) and torch.equal(r, rc)     # each step's results survived the later steps
        assert not torch.equal(outs[0][0]["agent_pos"], outs[2][0]["agent_pos"])
        assert outs[0][0]["agent_pos"].data_ptr() != outs[1][0]["agent_pos"].data_ptr()
        env.close()
        env = gym.make("gym_genesis/CubePick-v0", robot="franka", num_envs=n, host_io=host_io, alias_outputs=True).unwrapped
        env.reset(seed=1)
        o1 = env.step(torch.as_tensor(acts[0]).pin_memory() if host_io else acts[0])[0]
        o2 = env.step(torch.as_tensor(acts[1]).pin_memory() if host_io else acts[1])[0]
        assert o1["agent_pos"].data_ptr() == o2["agent_pos"].data_ptr()                # same buffers, overwritten in place
        env.close()


def test_host_io_masked_reset_push_and_control_after_host_step():
    """after a host-fed step the held PD targets live ON THE DEVICE (written by the kernel from the batch it consumed), so a
    masked reset, push() and control_dofs_position work, and push() re-applies the consumed action even if the caller has since
    overwritten its pinned buffer."""
    import gym_genesis  # noqa: F401
    from gym_genesis_b200.gym_api import gym
    n = 96
    rng = np.random.default_rng(2)
    env = gym.make("gym_genesis/CubePick-v0", robot="franka", num_envs=n, host_io=True, host_obs=True).unwrapped
    ref = gym.make("gym_genesis/CubePick-v0", robot="franka", num_envs=n).unwrapped
    env.reset(seed=5); ref.reset(seed=5)
    a = torch.as_tensor(f32(mdl.FRANKA_HOME + rng.uniform(-0.3, 0.3, (n, 9)))).pin_memory()
    obs, reward, term, trunc, info = env.step(a)
    ref.step(a.numpy().copy())
    assert obs["agent_pos"].device.type == "cpu" and obs["environment_state"].shape == (n, 11)      # host_obs: observations on the host
    assert torch.equal(obs["agent_pos"], ref._env.sim.obs_agent.cpu()) and torch.equal(obs["environment_state"], ref._env.sim.obs_env.cpu())
    assert env._env._targets.is_cuda and torch.equal(env._env._targets.cpu(), a)
    a.zero_()                                                   # the caller recycles its pinned buffer ...
    env.push(); ref.push()                                      # ... push() still holds the targets the step consumed
    for x, y in zip(env._env.sim.get_state(), ref._env.sim.get_state()):
        assert torch.equal(x, y)
    mask = torch.zeros(n, dtype=torch.bool); mask[::3] = True
    env.reset(options={"reset_mask": mask}); ref.reset(options={"reset_mask": mask})
    assert torch.equal(env._env._targets, ref._env._targets)
    env.get_robot().control_dofs_position(np.full((n, 2), 0.01, np.float32), [7, 8])
    ref.get_robot().control_dofs_position(np.full((n, 2), 0.01, np.float32), [7, 8])
    env.push(); ref.push()
    for x, y in zip(env._env.sim.get_state(), ref._env.sim.get_state()):
        assert torch.equal(x, y)
    env.close(); ref.close()


def test_fresh_host_results_and_observations_without_steady_state_allocation():
    """Default (fresh) outputs on the host-I/O path: reward / flags / host observations of a step keep their values after later steps,
    and a loop that drops its results recycles a bounded set of slabs (sim.py: _SlabRing) instead of allocating."""
    import gym_genesis  # noqa: F401
    from gym_genesis_b200.gym_api import gym
    n = 64
    rng = np.random.default_rng(4)
    env = gym.make("gym_genesis/CubePick-v0", robot="franka", num_envs=n, host_io=True, host_obs=True).unwrapped
    env.reset(seed=2)
    acts = [torch.as_tensor(f32(mdl.FRANKA_HOME + rng.uniform(-0.3, 0.3, (n, 9)))).pin_memory() for _ in range(3)]
    held = []
    for a in acts:
        obs, reward, term, trunc, info = env.step(a)
        assert obs["agent_pos"].device.type == "cpu" and obs["agent_pos"].is_pinned()
        held.append((obs, {k: v.clone() for k, v in obs.items()}, reward, reward.clone()))
    for obs, oc, r, rc in held:
        assert all(torch.equal(obs[k], oc[k]) for k in obs) and torch.equal(r, rc)
    assert not torch.equal(held[0][0]["agent_pos"], held[2][0]["agent_pos"])
    sim = env._env.sim
    del held, obs, reward, term, trunc
    K = sim._pool_size()
    for k in range(5 * K):
        env.step(acts[k % 3])
    rings = [sim._out_ring, sim._host_ring, sim._host_obs_ring]
    assert all(len(r.slabs) <= 3 for r in rings), [len(r.slabs) for r in rings]
    ids = [[id(s[0][0]) for s in r.slabs] for r in rings]
    for k in range(5 * K):
        env.step(acts[k % 3])
    assert ids == [[id(s[0][0]) for s in r.slabs] for r in rings]           # the same slabs again: nothing was allocated
    env.close()


def test_calls_restore_the_callers_current_device(Sim):
    """a simulator on another GPU must not change the caller's current CUDA device (torch reads it from the runtime)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    torch.cuda.set_device(0)
    m = mdl.franka_cube_pick_model()
    sim = Sim(m, 64, device="cuda:1")
    assert torch.cuda.current_device() == 0
    sim.reset()
    sim.step(torch.as_tensor(np.tile(f32(m.home_q), (64, 1)), device="cuda:1"))
    sim.step_host(torch.as_tensor(np.tile(f32(m.home_q), (64, 1))).pin_memory())
    assert torch.cuda.current_device() == 0
    assert torch.empty(1, device="cuda").device.index == 0
    assert sim.obs_agent.device.index == 1 and torch.isfinite(sim.obs_agent).all()
    sim.close()
    assert torch.cuda.current_device() == 0


def test_action_staging_paths_agree(Sim):
    """The block's action slab is staged with one bulk copy when the batch pointer is 16-byte aligned and the block is full; an unaligned
    batch (or the partial tail block) reads its rows directly.  Same results bit for bit, device-resident and pinned host batches."""
    m = mdl.franka_cube_pick_model()
    n = 32 * 7 + 13                                               # seven full blocks + a partial tail
    rng = np.random.default_rng(6)
    act = f32(m.home_q + rng.uniform(-0.3, 0.3, (n, 9)))
    outs = []
    for variant in ("aligned", "unaligned", "host_aligned", "host_unaligned"):
        sim = Sim(m, n, seed=9)
        sim.reset(None, None)
        flat = torch.zeros(n * 9 + 8, dtype=torch.float32, device="cpu" if variant.startswith("host") else "cuda")
        if variant.startswith("host"):
            flat = flat.pin_memory()
        off = 0 if variant.endswith("_aligned") or variant == "aligned" else 1
        a = flat[off:off + n * 9].view(n, 9)
        a.copy_(torch.as_tensor(act))
        assert (a.data_ptr() % 16 == 0) == (off == 0)
        for _ in range(3):
            if variant.startswith("host"):
                sim.step_host(a)
            else:
                sim.step(a)
        outs.append([t.clone() for t in sim.get_state()] + [sim.obs_agent.clone(), sim.obs_env.clone()])
        sim.close()
    for other in outs[1:]:
        for x, y in zip(outs[0], other):
            assert torch.equal(x, y)
