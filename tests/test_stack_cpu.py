"""CPU tier of the multi-object (CubeStack) path, SURVEY.md 8f-1: the oracle's stacking scene against closed forms and
conservation laws, the golden trajectory against the oracle that generated it, and the DEVICE pipeline of gg_stack.cuh --
including the warp-cooperative solve, emulated with 32 host threads per warp -- against the oracle."""
import ctypes as C
import os

import numpy as np
import pytest

from gym_genesis_b200 import cabi
from gym_genesis_b200.model import franka_cube_stack_model
from helpers import (TOL_Q_ABS, _p, f32, hostsim_stack_step, stack_cluster_states, stack_clutter_states, stack_pile_states,
                     stack_random_states, stack_substep_errors)

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
def gg(sm):
    return cabi.compile_gg(sm), cabi.compile_stack_task(sm)


def _rest_state(orc, sm, xy):
    q, qd, qa = orc.new_state(1)
    orc.reset(q, qd, qa, np.asarray(xy, dtype=np.float64)[None])
    return q, qd, qa


SPREAD = [[0.55, 0.05], [0.5, -0.15], [0.35, 0.25], [0.65, 0.25], [0.4, -0.25]]


# ---- the oracle's stacking scene -----------------------------------------------------------------------------------
def test_model_sizes_and_reference_constants(sm):
    assert (sm.nu, sm.n_obj, sm.nv, sm.nq) == (9, 5, 39, 44)
    # franka/cube_stack.py:125-126, 132-133, 141-142
    assert sm.obj_spawn[0] == (0.45, 0.75, -0.2, 0.2) and all(s == (0.3, 0.7, -0.3, 0.3) for s in sm.obj_spawn[1:])
    t = sm.dof_table()
    assert list(t["kp"][:9]) == [4500, 4500, 3500, 3500, 2000, 2000, 2000, 100, 100]       # :156
    assert list(t["kv"][:9]) == [450, 450, 350, 350, 200, 200, 200, 10, 10]                # :157
    assert list(t["frc_hi"][:9]) == [87] * 7 + [100, 100]                                  # :158-161
    assert (sm.stack_xy, sm.stack_z) == (0.05, 0.03)                                       # :195


def test_stacked_cube_rests_and_scores(orc, sm):
    xy = [list(x) for x in SPREAD]
    xy[0] = xy[1]
    q, qd, qa = _rest_state(orc, sm, xy)
    q[0, 9 + 2] = 0.06
    tgt = sm.home_q[None].copy()
    for _ in range(150):
        info, _ = orc.step(q, qd, qa, tgt)
    assert info[0, 0] == 20                                   # 4 x 4 cube-plane corners + 4 face contacts cube_1 on cube_2
    assert np.abs(qd[0, 9:21]).max() < 1e-6
    # closed form of the soft contact at rest: 4 corners share m g, each through 4 pyramidal rows with
    # D = 1 / (2 mu^2 (1 + mu^2) (1 - imp) / imp w), imp = dmax beyond the 1 mm width; top cube rests deeper on its support
    z1, z2 = q[0, 11], q[0, 18]
    assert 0.0195 < z2 < 0.02 and 0.039 < z1 - z2 < 0.04
    _, envs, rew, term = orc.obs(q)
    assert rew[0] == 1.0 and term[0]
    assert np.allclose(envs[0, 11:14], q[0, 16:19]) and np.allclose(envs[0, 0:3], q[0, 9:12])


def test_overlapping_spawn_is_pushed_apart(orc, sm):
    xy = [list(x) for x in SPREAD]
    xy[1] = [xy[0][0] + 0.02, xy[0][1] + 0.01]
    q, qd, qa = _rest_state(orc, sm, xy)
    assert len(orc.contacts(q[0])) > 20                       # cube-cube box-box contacts exist (box_box_detection=True)
    tgt = sm.home_q[None].copy()
    for _ in range(200):
        orc.step(q, qd, qa, tgt)
    d = q[0, 9:11] - q[0, 16:18]
    assert max(abs(d[0]), abs(d[1])) > 0.0399                 # separated along a face normal
    assert len(orc.contacts(q[0])) == 20 and np.all(np.isfinite(q))


def test_cube_cube_contact_forces_are_internal(orc, sm):
    """Two cubes colliding in free flight: the contact rows enter with J_b - J_a, so the pair's total linear momentum
    changes by gravity alone."""
    q, qd, qa = _rest_state(orc, sm, SPREAD)
    q[0, 9:12] = [0.5, 0.0, 0.5]
    q[0, 16:19] = [0.5, 0.036, 0.51]
    qd[0, 9:15] = [0, 0.4, 0, 0.3, -0.2, 0.1]
    qd[0, 15:21] = [0, -0.3, 0, 0, 0.5, 0]
    p0 = qd[0, 9:12] + qd[0, 15:18]
    info, _ = orc.step(q, qd, qa, sm.home_q[None].copy())
    assert info[0, 0] > 12                                    # 3 resting distractors x 4 corners + the cube-cube points
    p1 = qd[0, 9:12] + qd[0, 15:18]
    assert np.allclose(p1 - p0, [0, 0, -2 * 9.81 * 0.01], atol=1e-9)
    assert abs(qd[0, 10] - 0.4) > 1e-3                        # and the contact did act


def test_reward_thresholds_are_float32_strict(orc, sm):
    q, _, _ = _rest_state(orc, sm, SPREAD)
    Q = np.repeat(q, 6, axis=0)
    Q[:, 16:19] = [0.5, 0.0, 0.02]
    c = np.float32
    cases = [(c(0.5) + c(0.04), 0.0, 0.06, True), (0.5, 0.0, c(0.02) + c(0.03), False),     # z diff == 0.03f: not >
             (0.5, 0.0, np.nextafter(c(0.02) + c(0.03), c(1)), True), (0.55, 0.0, 0.06, False),   # xy == 0.05: not <
             (0.5, 0.0499, 0.06, True), (0.5, 0.0, 0.04, False)]
    for i, (x, y, z, _) in enumerate(cases):
        Q[i, 9:12] = [x, y, z]
    _, _, rew, term = orc.obs(f32(Q).astype(np.float64))
    assert [bool(r) for r in rew] == [e for *_, e in cases] and list(term) == [e for *_, e in cases]


def test_golden_trajectory_is_the_oracles(orc):
    g = np.load(GOLDEN)
    assert g["final_reward"][0] == 1.0 and g["ncon"].max() >= 48 and g["ncon"].min() == 16
    idx = np.arange(0, len(g["q"]), 7)
    q, qd, qa = g["q"][idx].copy(), g["qd"][idx].copy(), g["qacc"][idx].copy()
    info, _ = orc.step(q, qd, qa, g["act"][idx], iterations=100, tol=1e-12)
    assert np.abs(q - g["qn"][idx]).max() < 1e-12 and np.abs(qd - g["qdn"][idx]).max() < 1e-10
    assert (info[:, 0] == g["ncon"][idx]).all()


# ---- device pipeline on the host harness ---------------------------------------------------------------------------
def test_struct_layout_matches_header(hostsim):
    assert hostsim.hs_sizeof_stack_task() == C.sizeof(cabi.GgStackTask)


def _asm(o, Q, QD, T):
    return np.array([np.abs(o.dynamics(Q[k], QD[k], T[k])["a_smooth"]).max() * 0.01 for k in range(Q.shape[0])])


def test_substep_random_states(hostsim, gg, sm, orc):
    rng = np.random.default_rng(11)
    Q, QD, T = stack_random_states(sm, rng, 16)
    QA = np.zeros_like(QD)
    qo, qdo, qao = Q.copy(), QD.copy(), QA.copy()
    info, _ = orc.step(qo, qdo, qao, T, iterations=100, tol=1e-12)
    q, qd, qa, hinfo = hostsim_stack_step(hostsim, gg[0], gg[1], Q, QD, QA, T)
    assert (hinfo[:, 0] == info[:, 0]).all()                  # identical contact counts (incl. cube-cube pairs)
    assert info[1, 0] > 20 and info[2, 0] == 20 and info[3, 0] == 16
    eq, ev = stack_substep_errors(Q, QD, qo, qdo, q, qd, _asm(orc, Q, QD, T))
    assert eq.max() < TOL_Q_ABS and ev.max() < 1.0, (eq.max(), ev.max())
    # warm start from the converged accelerations: same answer, at most as many iterations
    q2, qd2, _, hinfo2 = hostsim_stack_step(hostsim, gg[0], gg[1], Q, QD, qao, T)
    eq2, ev2 = stack_substep_errors(Q, QD, qo, qdo, q2, qd2, _asm(orc, Q, QD, T))
    assert ev2.max() < 1.0 and (hinfo2[:, 1] <= np.maximum(hinfo[:, 1], 1)).all()


def test_substep_along_golden_stack_trajectory(hostsim, gg, orc):
    g = np.load(GOLDEN)
    idx = np.arange(0, len(g["q"]), 4)
    Q, QD, QA, T = g["q"][idx], g["qd"][idx], g["qacc"][idx], g["act"][idx]
    q, qd, qa, hinfo = hostsim_stack_step(hostsim, gg[0], gg[1], Q, QD, QA, T)
    assert (hinfo[:, 0] == g["ncon"][idx]).all()
    assert hinfo[:, 2].min() == 0 and hinfo[:, 2].max() == 1  # both the decoupled and the coupled path are exercised
    eq, ev = stack_substep_errors(Q, QD, g["qn"][idx], g["qdn"][idx], q, qd, g["asm"][idx])
    assert eq.max() < TOL_Q_ABS and ev.max() < 1.0, (eq.max(), ev.max(), int(np.argmax(ev)))


def test_all_bodies_active_two_rows_per_lane(hostsim, gg, orc):
    Q, QD, QA, T = stack_cluster_states(np.load(GOLDEN))
    qo, qdo, qao = Q.copy(), QD.copy(), QA.copy()
    info, _ = orc.step(qo, qdo, qao, T, iterations=200, tol=1e-12)
    q, qd, qa, hinfo = hostsim_stack_step(hostsim, gg[0], gg[1], Q, QD, QA, T)
    assert (hinfo[:, 3] == 63).all()                          # robot + five cubes in the cooperative solve: 39 dofs > 32 lanes
    assert (hinfo[:, 0] == info[:, 0]).all() and info[:, 0].max() >= 48
    eq, ev = stack_substep_errors(Q, QD, qo, qdo, q, qd, _asm(orc, Q, QD, T))
    assert eq.max() < TOL_Q_ABS and ev.max() < 1.0, (eq.max(), ev.max())


def test_tumbling_pile_states(hostsim, gg, sm, orc):
    """Cubes dropped on one spot with random orientations and spins: edge-edge / vertex-face cube-cube contacts, piles of three."""
    Q, QD, QA, T = stack_pile_states(sm, orc, n=6)
    qo, qdo, qao = Q.copy(), QD.copy(), QA.copy()
    info, _ = orc.step(qo, qdo, qao, T, iterations=200, tol=1e-12)
    q, qd, qa, hinfo = hostsim_stack_step(hostsim, gg[0], gg[1], Q, QD, QA, T)
    assert (hinfo[:, 0] == info[:, 0]).all() and info[:, 0].max() >= 10
    eq, ev = stack_substep_errors(Q, QD, qo, qdo, q, qd, _asm(orc, Q, QD, T))
    assert eq.max() < TOL_Q_ABS and ev.max() < 1.0, (eq.max(), ev.max(), int(np.argmax(ev)))


def test_gripper_cube_clutter_states(hostsim, gg, sm, orc):
    """Cubes overlapping the fingers, each other and the plane at random poses (up to ~75 contacts)."""
    Q, QD, T = stack_clutter_states(sm, gg[0], np.random.default_rng(123), 40)
    QA = np.zeros_like(QD)
    qo, qdo, qao = Q.copy(), QD.copy(), QA.copy()
    info, _ = orc.step(qo, qdo, qao, T, iterations=300, tol=1e-12)
    q, qd, qa, hinfo = hostsim_stack_step(hostsim, gg[0], gg[1], Q, QD, QA, T)
    assert (hinfo[:, 0] == info[:, 0]).all() and info[:, 0].max() >= 40
    eq, ev = stack_substep_errors(Q, QD, qo, qdo, q, qd, _asm(orc, Q, QD, T))
    assert eq.max() < TOL_Q_ABS and ev.max() < 1.0, (eq.max(), ev.max(), int(np.argmax(ev)))


def test_observation_and_flags_bit_exact(hostsim, gg, sm, orc):
    g = np.load(GOLDEN)
    Q = f32(np.concatenate([g["q"][::5], g["final_q"]]))
    n = Q.shape[0]
    agent, envs = np.zeros((n, 9), np.float32), np.zeros((n, 14), np.float32)
    rew, term = np.zeros(n, np.float32), np.zeros(n, np.uint8)
    assert hostsim.hs_stack_observe(C.byref(gg[0]), C.byref(gg[1]), n, _p(Q), _p(agent), _p(envs), _p(rew), _p(term)) == 0
    a, e, r, t = orc.obs(Q.astype(np.float64))
    assert np.abs(agent - a).max() < 2e-6 and np.abs(envs - e).max() < 2e-6
    assert (rew == r).all() and (term.astype(bool) == t).all() and t.any() and not t.all()


def test_spawn_draws_bit_exact(hostsim, gg, orc):
    for seed, gid, epi in [(0, 0, 0), (7, 123456, 3), (2 ** 40 + 5, 16383, 17)]:
        xy = np.zeros((5, 2), np.float32)
        hostsim.hs_stack_spawn(C.byref(gg[1]), C.c_uint64(seed), C.c_uint32(gid), C.c_uint32(epi), _p(xy))
        for k in range(5):
            ref = orc.spawn_obj(seed, gid, epi, k)
            assert xy[k].tobytes() == ref.tobytes()
            lo, hi = gg[1].spawn_lo[k], gg[1].spawn_hi[k]
            assert lo[0] <= xy[k, 0] <= hi[0] and lo[1] <= xy[k, 1] <= hi[1]


# ---- SO-101 CubeStackBatch ----------------------------------------------------------------------
@pytest.fixture(scope="module")
def so():
    from gym_genesis_b200.model_so101 import so101_cube_stack_model
    from oracle.binding import Oracle
    m = so101_cube_stack_model()
    return m, Oracle(m, np.float64), cabi.compile_gg(m), cabi.compile_stack_task(m)


def test_so101_stack_model_reference_constants(so):
    m = so[0]
    assert (m.nu, m.n_obj, m.nv, m.nq) == (6, 5, 36, 41)
    assert m.obj_spawn[0] == (-0.3, -0.1, -0.1, 0.1) and m.obj_spawn[2] == (-0.35, 0.0, -0.2, 0.2)    # cube_stack_batch.py:72-75, 93-94
    assert m.stack_min_dist == 0.06 and m.agent_mode == "qpos" and m.settle_steps == 1                # :70, :176, :111
    assert abs(m.spawn_z - (0.7000312834978104 + 0.021)) < 1e-12 and tuple(m.spawn_quat) == (0, 0, 0, 1)   # :66-67
    assert np.allclose(m.home_q, np.deg2rad([0, -177, 165, 72, -83, 0]))                              # :106
    assert set(m.dof_table()["kp"][:6]) == {17.8}                                                     # so101_old_calib.xml:23


def _so101_states(m, rng, n):
    nu, K = m.nu, m.n_obj
    Q = np.zeros((n, nu + 7 * K)); QD = np.zeros((n, nu + 6 * K))
    Q[:, :nu] = m.home_q + rng.uniform(-0.15, 0.15, (n, nu))
    QD[:, :nu] = rng.normal(0, 0.3, (n, nu))
    for k, (xlo, xhi, ylo, yhi) in enumerate(m.obj_spawn):
        o = nu + 7 * k
        Q[:, o] = rng.uniform(xlo, xhi, n); Q[:, o + 1] = rng.uniform(ylo, yhi, n)
        Q[:, o + 2] = m.plane_z + 0.02 - rng.uniform(0, 2e-4, n)
        yaw = rng.uniform(-np.pi, np.pi, n)
        Q[:, o + 3] = np.cos(yaw / 2); Q[:, o + 6] = np.sin(yaw / 2)
    Q[1, nu + 7:nu + 9] = Q[1, nu:nu + 2] + [0.02, 0.011]
    Q[2, nu + 7:nu + 14] = [Q[2, nu], Q[2, nu + 1], m.plane_z + 0.0199, 1, 0, 0, 0]
    Q[2, nu + 2:nu + 7] = [m.plane_z + 0.0597, 1, 0, 0, 0]
    T = m.home_q + rng.uniform(-0.2, 0.2, (n, nu))
    return f32(Q).astype(np.float64), f32(QD).astype(np.float64), f32(T).astype(np.float64)


def test_so101_substep_and_observation(hostsim, so):
    m, orc, ggm, task = so
    rng = np.random.default_rng(5)
    Q, QD, T = _so101_states(m, rng, 12)
    QA = np.zeros_like(QD)
    qo, qdo, qao = Q.copy(), QD.copy(), QA.copy()
    info, _ = orc.step(qo, qdo, qao, T, iterations=100, tol=1e-12)
    q, qd, qa, hinfo = hostsim_stack_step(hostsim, ggm, task, Q, QD, QA, T)
    assert (hinfo[:, 0] == info[:, 0]).all() and info[1, 0] > 20
    eq, ev = stack_substep_errors(Q, QD, qo, qdo, q, qd, _asm(orc, Q, QD, T))
    assert eq.max() < TOL_Q_ABS and ev.max() < 1.0, (eq.max(), ev.max())
    n = Q.shape[0]
    agent, envs = np.zeros((n, 8), np.float32), np.zeros((n, 14), np.float32)
    rew, term = np.zeros(n, np.float32), np.zeros(n, np.uint8)
    assert hostsim.hs_stack_observe(C.byref(ggm), C.byref(task), n, _p(f32(q)), _p(agent), _p(envs), _p(rew), _p(term)) == 0
    a, e, r, t = orc.obs(q.astype(np.float64))
    assert a.shape == (n, 6) and (agent[:, :6] == f32(q)[:, :6]).all() and np.abs(a - q[:, :6]).max() == 0    # agent_pos = qpos
    assert np.abs(envs - e).max() < 2e-6 and (rew == r).all() and (term.astype(bool) == t).all() and t[2]


def test_so101_spawn_rejection_bit_exact(hostsim, so):
    m, orc, ggm, task = so
    redrawn = 0
    for gid in range(400):
        xy = np.zeros((5, 2), np.float32)
        hostsim.hs_stack_spawn(C.byref(task), C.c_uint64(9), C.c_uint32(gid), C.c_uint32(2), _p(xy))
        ref = np.array([orc.spawn_obj(9, gid, 2, k) for k in range(5)])
        assert xy.tobytes() == ref.tobytes()
        assert np.hypot(*(xy[0] - xy[1])) >= 0.06 - 1e-6                                  # cube_stack_batch.py:70-86
        first = np.zeros((5, 2), np.float32)
        t0 = cabi.GgStackTask.from_buffer_copy(task); t0.min_dist = 0.0
        hostsim.hs_stack_spawn(C.byref(t0), C.c_uint64(9), C.c_uint32(gid), C.c_uint32(2), _p(first))
        redrawn += int((first[:2] != xy[:2]).any())
        assert (first[2:] == xy[2:]).all()
    assert 5 < redrawn < 120            # P(|d| < 0.06) on a 0.2 x 0.2 square is ~ 0.24


def test_cube_cube_broad_phase_is_conservative(hostsim, gg, sm, orc):
    """boxes_may_touch (bounding spheres + the six face axes of the SAT) may only err towards "may touch": whenever the oracle's
    box-box narrow phase finds a contact between two cubes, the flag is set; and it rejects most near misses."""
    rng = np.random.default_rng(8)
    n = 6000
    A = np.zeros((n, 7)); B = np.zeros((n, 7))
    A[:, :3] = [0.5, 0.4, 0.5]                                            # in the air, away from the gripper
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    B[:, :3] = A[:, :3] + d * rng.uniform(0.030, 0.072, (n, 1))          # centre distance from deep overlap to beyond the spheres
    for X in (A, B):
        q = rng.normal(size=(n, 4)); X[:, 3:] = q / np.linalg.norm(q, axis=1, keepdims=True)
    A[::4, 3:] = [1, 0, 0, 0]; B[::4, 3:] = [1, 0, 0, 0]                  # axis-aligned pairs: resting neighbours
    A, B = f32(A), f32(B)
    flag = np.zeros(n, np.uint8)
    hostsim.hs_cubes_may_touch(C.byref(gg[0]), n, _p(A), _p(B), _p(flag))
    q0, _, _ = orc.new_state(1)
    q0[0, :9] = sm.home_q
    for k in range(2, 5):
        q0[0, 9 + 7 * k: 9 + 7 * k + 7] = [2.0 + k, 0, 0.5, 1, 0, 0, 0]   # the other cubes far away, in the air
    touching = np.zeros(n, bool)
    for i in range(n):
        q0[0, 9:16] = A[i]; q0[0, 16:23] = B[i]
        c = orc.contacts(q0[0])
        touching[i] = any({int(r[7]), int(r[8])} == {sm.obj_bodies[0], sm.obj_bodies[1]} for r in c)
    assert touching.sum() > 1000 and (~touching).sum() > 1000
    assert flag[touching].all(), int((~flag[touching].astype(bool)).sum())
    assert flag[~touching].mean() < 0.5                                   # the nine edge-edge axes are all it leaves on the table


def test_kitchen_model_reference_constants_and_substep(hostsim):
    """FrankaCubeStackKitchenBatch: Panda at morph scale 0.6 at (-0.5, 0, 0.7) (tasks/utils.py:369-376), cubes on the island top,
    spawn ranges of cube_stack_kitchen_batch.py:71-88; the host build of the device pipeline against the oracle on its states."""
    from gym_genesis_b200.model import KITCHEN_ISLAND_TOP_Z, franka_cube_pick_model, franka_cube_stack_kitchen_model
    from helpers import stack_clutter_states
    from oracle.binding import Oracle
    k, full = franka_cube_stack_kitchen_model(), franka_cube_pick_model()
    assert k.obj_spawn[0] == (-0.3, -0.1, -0.15, 0.15) and k.obj_spawn[2] == (-0.35, 0.0, -0.2, 0.2)
    assert abs(k.plane_z - 0.7000312834978104) < 1e-12 and abs(k.spawn_z - (KITCHEN_ISLAND_TOP_Z + 0.021)) < 1e-12
    assert tuple(k.bodies[0].pos) == (-0.5, 0.0, 0.7)
    for b6, b1 in zip(k.bodies[:11], full.bodies[:11]):
        assert np.allclose(b6.mass, b1.mass * 0.6 ** 3) and np.allclose(b6.inertia, np.asarray(b1.inertia) * 0.6 ** 5)
        if b6.name != "link0":
            assert np.allclose(b6.pos, np.asarray(b1.pos) * 0.6)
    assert list(k.dof_table()["kp"][:9]) == [4500, 4500, 3500, 3500, 2000, 2000, 2000, 100, 100]      # :101
    xp, _ = k.fk(np.concatenate([k.home_q, k.default_cube_q()]))
    xf, _ = full.fk(full.full_q(full.home_q))
    assert np.allclose(xp[8] - [-0.5, 0.0, 0.7], 0.6 * xf[8], atol=1e-9)                                # the scaled arm, same pose
    gm, task = cabi.compile_gg(k), cabi.compile_stack_task(k)
    orc = Oracle(k, np.float64)
    Q, QD, T = stack_clutter_states(k, gm, np.random.default_rng(61), 24)
    QA = np.zeros_like(QD)
    qo, qdo, qao = Q.copy(), QD.copy(), QA.copy()
    info, _ = orc.step(qo, qdo, qao, T, iterations=300, tol=1e-12)
    q, qd, qa, hinfo = hostsim_stack_step(hostsim, gm, task, Q, QD, QA, T)
    assert (hinfo[:, 0] == info[:, 0]).all()
    eq, ev = stack_substep_errors(Q, QD, qo, qdo, q, qd, _asm(orc, Q, QD, T))
    assert eq.max() < TOL_Q_ABS and ev.max() < 1.0, (eq.max(), ev.max())
