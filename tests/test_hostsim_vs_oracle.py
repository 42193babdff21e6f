"""CPU tier: the per-env DEVICE pipeline (gg_env.cuh, compiled for the host as a test harness) against the
oracle.  Catches kernel-math bugs without a GPU; the GPU tier repeats these through the C ABI."""
import ctypes as C

import numpy as np
import pytest

from gym_genesis_b200 import cabi
from gym_genesis_b200 import model as mdl
from helpers import (TOL_Q_ABS, _p, f32, hostsim_step, pick_clutter_states, random_states, scripted_grasp_states, substep_errors)


@pytest.fixture(scope="module")
def gg(franka_model):
    return cabi.compile_gg(franka_model)


def test_struct_layout_matches_header(hostsim):
    assert hostsim.hs_sizeof_model() == C.sizeof(cabi.GgModel)


def test_robot_dynamics(hostsim, gg, franka_model, oracle64):
    rng = np.random.default_rng(3)
    Q, QD, T = random_states(franka_model, rng, 8, cube_on_ground=False)
    for k in range(8):
        d = oracle64.dynamics(Q[k], QD[k], T[k])
        Mt, fs = np.zeros((9, 9), np.float32), np.zeros(9, np.float32)
        hostsim.hs_dynamics(C.byref(gg), _p(f32(Q[k])), _p(f32(QD[k])), _p(f32(T[k])), _p(Mt), _p(fs))
        assert np.abs(Mt - d["Mt"][:9, :9]).max() < 5e-6
        fs_ref = (d["tau"] - d["bias"])[:9]
        assert np.abs(fs - fs_ref).max() < 1e-4 * max(1.0, np.abs(fs_ref).max())


def _a_smooth_scale(o, Q, QD, T):
    return np.array([np.abs(o.dynamics(Q[k], QD[k], T[k])["a_smooth"]).max() * 0.01 for k in range(Q.shape[0])])


@pytest.mark.parametrize("on_ground", [True, False])
def test_substep_random_states(hostsim, gg, franka_model, oracle64, on_ground):
    rng = np.random.default_rng(4 + on_ground)
    Q, QD, T = random_states(franka_model, rng, 64, cube_on_ground=on_ground)
    QA = np.zeros_like(QD)
    qo, qdo, qao = Q.copy(), QD.copy(), QA.copy()
    info, _ = oracle64.step(qo, qdo, qao, T, iterations=100, tol=1e-12)
    q, qd, qa, hinfo, _ = hostsim_step(hostsim, gg, Q, QD, QA, T)
    assert (hinfo[:, 0] == info[:, 0]).all()                    # identical contact counts
    eq, ev = substep_errors(Q, QD, qo, qdo, q, qd, _a_smooth_scale(oracle64, Q, QD, T))
    assert eq.max() < TOL_Q_ABS and ev.max() < 1.0, (eq.max(), ev.max())


def test_substep_along_scripted_grasp(hostsim, gg, franka_model, oracle64):
    states, reward = scripted_grasp_states(franka_model, oracle64, stride=3)
    assert reward == 1.0
    worst_q = worst_v = 0.0
    seen_heavy = False
    for Q, QD, QA, T in states:
        Q, QD, QA, T = (f32(x).astype(np.float64) for x in (Q, QD, QA, T))
        qo, qdo, qao = Q.copy(), QD.copy(), QA.copy()
        info, _ = oracle64.step(qo, qdo, qao, T, iterations=100, tol=1e-12)
        q, qd, qa, hinfo, _ = hostsim_step(hostsim, gg, Q, QD, QA, T)
        assert hinfo[0, 0] == info[0, 0]
        seen_heavy |= info[0, 0] >= 32
        eq, ev = substep_errors(Q, QD, qo, qdo, q, qd, _a_smooth_scale(oracle64, Q, QD, T))
        worst_q, worst_v = max(worst_q, eq.max()), max(worst_v, ev.max())
    assert seen_heavy                                            # pad-cube and pad-plane manifolds were exercised
    assert worst_q < TOL_Q_ABS and worst_v < 1.0, (worst_q, worst_v)


def test_observation_and_flags(hostsim, gg, franka_model, oracle64):
    rng = np.random.default_rng(6)
    Q, _, _ = random_states(franka_model, rng, 32, cube_on_ground=False)
    Q[:8, 11] = np.float32(0.1)            # exactly on the strict threshold: z > 0.1f is False
    Q[8:16, 11] = np.nextafter(np.float32(0.1), np.float32(1.0))
    a, e, r, t = oracle64.obs(Q)
    agent, envs = np.zeros((32, 9), np.float32), np.zeros((32, 11), np.float32)
    rew, term = np.zeros(32, np.float32), np.zeros(32, np.uint8)
    hostsim.hs_observe(C.byref(gg), 32, _p(f32(Q)), _p(agent), _p(envs), _p(rew), _p(term))
    assert np.abs(agent - a).max() < 2e-6 and np.abs(envs - e).max() < 2e-6
    assert (term.astype(bool) == t).all() and (rew == r).all()            # bit-exact flags
    assert not t[:8].any() and t[8:16].all()


def test_spawn_draw_bit_exact(hostsim, gg, oracle64):
    for seed, gid, epi in [(0, 0, 0), (42, 7, 3), (2 ** 40 + 5, 131071, 200), (2 ** 64 - 1, 2 ** 31, 1)]:
        xy = np.zeros(2, np.float32)
        hostsim.hs_spawn(C.byref(gg), C.c_uint64(seed), C.c_uint32(gid), C.c_uint32(epi), _p(xy))
        ref = oracle64.spawn_xy(seed, gid, epi)
        assert xy.tobytes() == ref.tobytes()
        assert 0.45 <= xy[0] <= 0.80 and -0.25 <= xy[1] <= 0.25


def test_so101_substep_random_states(hostsim, so101_model, so101_oracle64):
    """Second articulation (5 hinges + hinge jaw, scaled, yawed base, plane at the island top, mu = 5)."""
    gg101 = cabi.compile_gg(so101_model)
    assert (gg101.n_arm, gg101.n_fing, gg101.fing_hinge, gg101.n_pads) == (5, 1, 1, 2)
    rng = np.random.default_rng(8)
    for on_ground in (True, False):
        Q, QD, T = random_states(so101_model, rng, 48, cube_on_ground=on_ground)
        QA = np.zeros_like(QD)
        qo, qdo, qao = Q.copy(), QD.copy(), QA.copy()
        info, _ = so101_oracle64.step(qo, qdo, qao, T, iterations=100, tol=1e-12)
        q, qd, qa, hinfo, _ = hostsim_step(hostsim, gg101, Q, QD, QA, T)
        assert (hinfo[:, 0] == info[:, 0]).all()
        sc = np.array([np.abs(so101_oracle64.dynamics(Q[k], QD[k], T[k])["a_smooth"]).max() * 0.01 for k in range(Q.shape[0])])
        eq, ev = substep_errors(Q, QD, qo, qdo, q, qd, sc)
        assert eq.max() < TOL_Q_ABS and ev.max() < 1.0, (on_ground, eq.max(), ev.max())


def test_so101_jaw_contacts_and_observation(hostsim, so101_model, so101_oracle64):
    """Drive the cube between the jaws so that the box-box manifolds of the jaw boxes are exercised."""
    m, o = so101_model, so101_oracle64
    gg101 = cabi.compile_gg(m)
    q0 = m.full_q(m.home_q)
    xp, xm = m.fk(q0)
    # put the cube where the fixed-jaw box is (guaranteed overlap) and a little above the plane
    g = m.geoms[0]
    c = xp[g.body] + xm[g.body] @ np.asarray(g.pos)
    Q = np.tile(q0, (4, 1))
    Q[:, 6:9] = c + np.array([[0.0, 0, 0], [0.01, 0, 0], [0, 0.01, 0.005], [0.0, -0.01, 0.01]])
    Q = f32(Q).astype(np.float64)
    QD = np.zeros((4, 12)); QA = np.zeros((4, 12)); T = np.tile(m.home_q, (4, 1))
    qo, qdo, qao = Q.copy(), QD.copy(), QA.copy()
    info, _ = o.step(qo, qdo, qao, T, iterations=200, tol=1e-12)
    assert info[:, 0].min() >= 1
    q, qd, qa, hinfo, _ = hostsim_step(hostsim, gg101, Q, QD, QA, T, iters=50)
    assert (hinfo[:, 0] == info[:, 0]).all()
    assert np.abs(q - qo).max() < 1e-4 and np.abs(qd - qdo).max() < 2e-2 * max(1.0, np.abs(qdo).max())
    a, e, r, t = o.obs(qo)
    agent, envs = np.zeros((4, 8), np.float32), np.zeros((4, 11), np.float32)
    rew, term = np.zeros(4, np.float32), np.zeros(4, np.uint8)
    hostsim.hs_observe(C.byref(gg101), 4, _p(f32(qo)), _p(agent), _p(envs), _p(rew), _p(term))
    assert np.abs(agent - a).max() < 2e-6 and np.abs(envs - e).max() < 2e-6 and (term.astype(bool) == t).all()


def test_scripted_grasp_rollout_open_loop(hostsim, gg, franka_model, oracle64):
    """200-step rollouts of the reference expert's stage machine (tests/golden/grasp_traj.npz, recorded closed-loop in
    the fp64 oracle) replayed open-loop through the device pipeline: contact-free phases stay at fp32 noise, the
    100-step mark is inside the stated rollout bound, success flags agree."""
    import os
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "grasp_traj.npz"))
    K = d["xy"].shape[0]
    Q, QD, QA = oracle64.new_state(K)
    oracle64.reset(Q, QD, QA, d["xy"].astype(np.float64))
    q, qd, qa, _, _ = hostsim_step(hostsim, gg, Q, QD, QA, np.tile(f32(mdl.FRANKA_HOME), (K, 1)))
    succ = np.zeros(K, bool)
    seen = 0
    for t in range(200):
        q, qd, qa, info, _ = hostsim_step(hostsim, gg, q, qd, qa, d["actions"][:, t])
        succ |= q[:, 11] > np.float32(0.1)
        seen = max(seen, info[:, 0].max())
        er = np.abs(q[:, :9] - d["qrob"][:, t]).max(1)
        ec = np.abs(q[:, 9:12] - d["cube"][:, t]).max(1)
        if t < 80:
            assert er.max() < 5e-6 and ec.max() < 1e-6, (t, er.max(), ec.max())
        if t == 99:
            assert er.max() < 1e-2 and ec.max() < 1e-2, (er.max(), ec.max())          # stated 100-step bound (helpers.ROLLOUT_POS_TOL)
            assert np.median(ec) < 1e-5
    assert seen >= 48
    assert (ec < 5e-3).mean() >= 0.9                                                   # chaotic contact phases: most still agree at step 200
    ref = d["reward"].max(1) > 0
    assert (succ == ref).mean() >= 0.9 and 0.2 < ref.mean() < 0.8


@pytest.mark.parametrize("robot", ["franka", "so101"])
def test_pad_broad_phase_is_conservative(hostsim, robot, franka_model, so101_model, oracle64, so101_oracle64):
    """The body-level broad phase (bounding spheres refined by sphere-vs-box, six face axes of the SAT and the exact box-plane reach)
    may only err towards "may touch": whenever the oracle's narrow phase finds a pad contact, the flag is set.  Cubes are placed at
    random poses within a few centimetres of the gripper, where the refinements actually decide."""
    m, orc = (franka_model, oracle64) if robot == "franka" else (so101_model, so101_oracle64)
    gg = cabi.compile_gg(m)
    rng = np.random.default_rng(12)
    n, nu = 3000, m.nu
    Q = np.zeros((n, nu + 7))
    Q[:, :nu] = m.home_q + rng.uniform(-0.5, 0.5, (n, nu))
    lo, hi = np.array(list(gg.lim_lo)[:nu]), np.array(list(gg.lim_hi)[:nu])
    Q[:, :nu] = np.clip(Q[:, :nu], lo, hi)
    pad_geoms = [g for g in m.geoms if g.group == 0]
    for k in range(n):
        xp, xm = m.fk(m.full_q(Q[k, :nu]))
        g = pad_geoms[rng.integers(len(pad_geoms))]
        tip = xp[g.body] + xm[g.body] @ np.asarray(g.pos)
        Q[k, nu:nu + 3] = tip + rng.uniform(-0.045, 0.045, 3)
        if k % 3 == 0:
            Q[k, nu + 2] = max(Q[k, nu + 2], m.plane_z + 0.0199)          # resting height: pad-plane and pad-cube both in play
        qq = rng.normal(size=4)
        Q[k, nu + 3:nu + 7] = qq / np.linalg.norm(qq)
    Q = f32(Q).astype(np.float64)
    flag = np.zeros(n, np.uint8)
    hostsim.hs_pads_may_touch(C.byref(gg), n, _p(f32(Q)), _p(flag))
    pad_bodies = {g.body for g in pad_geoms}
    touching = np.zeros(n, bool)
    for k in range(n):
        c = orc.contacts(Q[k])
        touching[k] = any(int(r[7]) in pad_bodies or int(r[8]) in pad_bodies for r in c)
    assert touching.sum() > 100 and (~touching).sum() > 100
    assert flag[touching].all(), int((~flag[touching].astype(bool)).sum())
    # and it is a useful filter: a good share of the near misses is rejected
    assert flag[~touching].mean() < 0.8


def test_substep_gripper_cube_clutter(hostsim, gg, franka_model, oracle64):
    """The cube at random poses around the fingers (overlapping pads and / or plane): contact counts and the substep against the oracle."""
    Q, QD, T = pick_clutter_states(franka_model, gg, np.random.default_rng(55), 300)
    QA = np.zeros_like(QD)
    qo, qdo, qao = Q.copy(), QD.copy(), QA.copy()
    info, _ = oracle64.step(qo, qdo, qao, T, iterations=300, tol=1e-12)
    q, qd, qa, hinfo, _ = hostsim_step(hostsim, gg, Q, QD, QA, T)
    assert (hinfo[:, 0] == info[:, 0]).all() and (info[:, 0] > 4).sum() > 50
    eq, ev = substep_errors(Q, QD, qo, qdo, q, qd, _a_smooth_scale(oracle64, Q, QD, T))
    assert eq.max() < TOL_Q_ABS and ev.max() < 1.0, (eq.max(), ev.max(), int(np.argmax(ev)))


@pytest.mark.parametrize("n_beyond", [1, 2, 3, 5])
def test_robot_block_with_k_joints_at_their_limits(hostsim, gg, franka_model, oracle64, n_beyond):
    """The robot block's joint-limit rows: the closed-form active-set solve (<= 2 active rows, gg_env.cuh solve_robot_limits_direct)
    and its Newton fallback (more rows) against the converged fp64 oracle.  k joints start beyond their limits, moving outwards, with
    targets that push further out; the others are random."""
    rng = np.random.default_rng(20 + n_beyond)
    n = 48
    Q, QD, T = random_states(franka_model, rng, n, cube_on_ground=False)
    lo, hi = np.array(list(gg.lim_lo)[:9]), np.array(list(gg.lim_hi)[:9])
    for e in range(n):
        for d in rng.choice(9, n_beyond, replace=False):
            up = rng.random() < 0.5
            Q[e, d] = (hi[d] + rng.uniform(1e-4, 3e-3)) if up else (lo[d] - rng.uniform(1e-4, 3e-3))
            QD[e, d] = rng.uniform(0.0, 0.3) * (1 if up else -1)
            T[e, d] = Q[e, d] + (0.2 if up else -0.2) * (0.02 if d >= 7 else 1.0)
    Q, QD, T = (f32(x).astype(np.float64) for x in (Q, QD, T))
    QA = np.zeros_like(QD)
    qo, qdo, qao = Q.copy(), QD.copy(), QA.copy()
    oracle64.step(qo, qdo, qao, T, iterations=200, tol=1e-12)
    q, qd, qa, hinfo, _ = hostsim_step(hostsim, gg, Q, QD, QA, T)
    eq, ev = substep_errors(Q, QD, qo, qdo, q, qd, _a_smooth_scale(oracle64, Q, QD, T))
    assert eq.max() < TOL_Q_ABS and ev.max() < 1.0, (eq.max(), ev.max())
    assert hinfo[:, 1].min() >= 1                                  # every env had limit rows to solve
