"""Shared helpers for the parity tests: scripted-grasp state generator, host-harness wrappers, tolerances."""
import ctypes as C

import numpy as np

from gym_genesis_b200.model import FRANKA_HOME, quat_to_mat

# ---- parity tolerances (stated once; see DESIGN.md "Parity") -----------------------------------
# Single substep, CUDA fp32 vs oracle fp64 run to convergence.  Velocity deltas are compared relative
# to the larger of the net delta and a floor: in contact-rich states the net delta is the small
# difference of large constraint/actuator terms and the fp32 penetration depth (|p| ~ 0.6 m, depth
# ~ 1e-4 m) carries ~1e-3 relative quantisation, which the stiffness K = 2770 1/s^2 turns into
# ~2e-4 m/s^2 of reference acceleration per row (the fp32 build of the oracle shows the same spread).
TOL_Q_ABS = 2e-6            # positions / quaternion components after one substep
TOL_DQD_REL = 1e-4          # |dqd - dqd_ref|_inf <= TOL_DQD_REL * max(|dqd_ref|_inf, |a_smooth|_inf * dt) + floor
TOL_DQD_FLOOR = 3e-5        # rad/s (= 3e-3 rad/s^2 * dt): fp32 geometry quantisation floor with ~50 contacts
ROLLOUT_POS_TOL = 1e-2      # 100-step closed-form-free rollout, free-space / light contact


def ik_numpy(m, target_pos, target_R, q0, iters=60):
    """Damped least-squares IK on the numpy model (stands in for robot.inverse_kinematics of the expert,
    reference examples/franka/pick_cube_state.py:46-51)."""
    q = q0.copy()
    for _ in range(iters):
        fq = m.full_q(q)
        xp, xm = m.fk(fq)
        ep = target_pos - xp[m.eef_body]
        Re = target_R @ xm[m.eef_body].T
        ang = np.arccos(np.clip((np.trace(Re) - 1) / 2, -1, 1))
        er = np.zeros(3) if ang < 1e-9 else ang / (2 * np.sin(ang)) * np.array([Re[2, 1] - Re[1, 2], Re[0, 2] - Re[2, 0], Re[1, 0] - Re[0, 1]])
        Jv, Jw = m.jacobian(fq, m.eef_body, xp[m.eef_body])
        J = np.vstack([Jv[:, :m.n_arm], Jw[:, :m.n_arm]])
        e = np.concatenate([ep, er])
        q[:m.n_arm] += J.T @ np.linalg.solve(J @ J.T + 1e-6 * np.eye(6), e)
        if np.linalg.norm(e) < 1e-10:
            break
    return q


def scripted_grasp_states(m, orc, cube_xy=(0.6, 0.1), stride=1):
    """Run the reference expert's stage machine (hover/stabilize/grasp/grasp/lift x 40 steps) in the oracle and
    return the visited (q, qd, qacc, action) tuples plus the final reward."""
    Rdown = quat_to_mat([0, 1, 0, 0])
    Q, QD, QA = orc.new_state(1)
    orc.reset(Q, QD, QA, np.array([cube_xy]))
    out = []
    act = FRANKA_HOME.copy()[None]
    out.append((Q.copy(), QD.copy(), QA.copy(), act.copy()))
    orc.step(Q, QD, QA, act)
    k = 0
    for dz, grip in [(0.115, 0.04), (0.115, 0.04), (0.03, -0.02), (0.03, -0.02), (0.25, -0.02)]:
        for _ in range(40):
            qt = ik_numpy(m, Q[0, 9:12] + np.array([0, 0, dz]), Rdown, Q[0, :9].copy())
            act = np.concatenate([qt[:7], [grip, grip]])[None]
            if k % stride == 0:
                out.append((Q.copy(), QD.copy(), QA.copy(), act.copy()))
            orc.step(Q, QD, QA, act, iterations=100, tol=1e-10)
            k += 1
    return out, orc.obs(Q)[2][0]


def f32(x):
    return np.ascontiguousarray(x, dtype=np.float32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def hostsim_step(hs, gg_model, Q, QD, QA, T, iters=10, tol=1e-5):
    q, qd, qa, t = f32(Q).copy(), f32(QD).copy(), f32(QA).copy(), f32(T)
    n = q.shape[0]
    info = np.zeros((n, 3), np.int32)
    grad = np.zeros(n, np.float32)
    hs.hs_step(C.byref(gg_model), n, _p(q), _p(qd), _p(qa), _p(t), iters, C.c_float(tol), _p(info), _p(grad))
    return q, qd, qa, info, grad


def random_states(m, rng, n, cube_on_ground=True):
    """Random robot configurations near home with random velocities; cube either resting or in flight."""
    nu = m.nu
    Q = np.zeros((n, nu + 7))
    QD = np.zeros((n, nu + 6))
    Q[:, :nu] = m.home_q + rng.uniform(-0.3, 0.3, (n, nu))
    fing_hi = 0.04 if m.name == "franka" else 1.0
    Q[:, m.n_arm:nu] = rng.uniform(0.0, fing_hi, (n, m.n_fing))
    Q[:, nu] = rng.uniform(m.spawn_lo[0], m.spawn_hi[0], n)
    Q[:, nu + 1] = rng.uniform(m.spawn_lo[1], m.spawn_hi[1], n)
    if cube_on_ground:
        Q[:, nu + 2] = m.plane_z + 0.02 - rng.uniform(0, 2e-4, n)
        yaw = rng.uniform(-np.pi, np.pi, n)
        Q[:, nu + 3] = np.cos(yaw / 2)
        Q[:, nu + 6] = np.sin(yaw / 2)
        QD[:, nu:nu + 2] = rng.normal(0, 0.05, (n, 2))
    else:
        Q[:, nu + 2] = m.plane_z + rng.uniform(0.1, 0.4, n)
        qq = rng.normal(size=(n, 4))
        Q[:, nu + 3:nu + 7] = qq / np.linalg.norm(qq, axis=1, keepdims=True)
        QD[:, nu:] = rng.normal(0, 0.5, (n, 6))
    QD[:, :nu] = rng.normal(0, 0.5, (n, nu))
    T = m.home_q + rng.uniform(-0.3, 0.3, (n, nu))
    T[:, m.n_arm:] = rng.uniform(-0.02, fing_hi, (n, m.n_fing))
    return f32(Q).astype(np.float64), f32(QD).astype(np.float64), f32(T).astype(np.float64)


def substep_errors(q0, qd0, q_ref, qd_ref, q, qd, a_smooth_scale):
    """(position error, velocity-delta error normalised by its bound: <= 1 passes).
    The position error is rescaled so that TOL_Q_ABS stays the threshold: positions integrate the new velocity over
    dt = 0.01 s, so their bound is TOL_Q_ABS + dt * (velocity bound)."""
    eq = np.abs(q - q_ref).max(axis=1)
    d_ref = np.abs(qd_ref - qd0).max(axis=1)
    ev = np.abs(qd - qd_ref).max(axis=1)
    bound = TOL_DQD_REL * np.maximum(d_ref, a_smooth_scale) + TOL_DQD_FLOOR
    return eq * TOL_Q_ABS / (TOL_Q_ABS + 0.01 * bound), ev / bound


# ---- multi-object (CubeStack) scene ---------------------------------------------------------------------------------
def hostsim_stack_step(hs, gg_model, task, Q, QD, QA, T, iters=10, tol=1e-5):
    """One substep of the multi-object pipeline (gg_stack.cuh) on the host harness; the warp-cooperative solve runs on 32
    host threads per emulated warp."""
    q, qd, qa, t = f32(Q).copy(), f32(QD).copy(), f32(QA).copy(), f32(T)
    n = q.shape[0]
    info = np.zeros((n, 4), np.int32)
    rc = hs.hs_stack_step(C.byref(gg_model), C.byref(task), n, _p(q), _p(qd), _p(qa), _p(t), iters, C.c_float(tol), _p(info))
    assert rc == 0
    return q, qd, qa, info


def stack_random_states(m, rng, n):
    """Reset-like states of the stacking scene: cubes on the plane at random spawn positions (with yaw and a small planar
    velocity), env 1 with an overlapping pair, env 2 with cube_1 resting on cube_2, env 3 with cube_1 in free flight;
    robot near home with random joint velocities."""
    nu, K = m.nu, m.n_obj
    Q = np.zeros((n, nu + 7 * K))
    QD = np.zeros((n, nu + 6 * K))
    Q[:, :nu] = m.home_q + rng.uniform(-0.25, 0.25, (n, nu))
    Q[:, m.n_arm:nu] = rng.uniform(0.0, 0.04, (n, m.n_fing))
    QD[:, :nu] = rng.normal(0, 0.4, (n, nu))
    for k, (xlo, xhi, ylo, yhi) in enumerate(m.obj_spawn):
        o = nu + 7 * k
        Q[:, o] = rng.uniform(xlo, xhi, n)
        Q[:, o + 1] = rng.uniform(ylo, yhi, n)
        Q[:, o + 2] = m.plane_z + 0.02 - rng.uniform(0, 2e-4, n)
        yaw = rng.uniform(-np.pi, np.pi, n)
        Q[:, o + 3] = np.cos(yaw / 2)
        Q[:, o + 6] = np.sin(yaw / 2)
        QD[:, nu + 6 * k:nu + 6 * k + 2] = rng.normal(0, 0.05, (n, 2))
    if n > 1:
        Q[1, nu + 7:nu + 9] = Q[1, nu:nu + 2] + [0.021, 0.012]                 # cube_2 overlaps cube_1
    if n > 2:
        Q[2, nu + 7:nu + 14] = [Q[2, nu], Q[2, nu + 1], m.plane_z + 0.0199, 1, 0, 0, 0]    # cube_1 on top of cube_2
        Q[2, nu + 2:nu + 7] = [m.plane_z + 0.0597, 1, 0, 0, 0]
        QD[2, nu:nu + 12] = 0
    if n > 3:
        Q[3, nu + 2] = m.plane_z + 0.3
        QD[3, nu:nu + 6] = rng.normal(0, 0.5, 6)
    T = m.home_q + rng.uniform(-0.3, 0.3, (n, nu))
    T[:, m.n_arm:] = rng.uniform(-0.02, 0.04, (n, m.n_fing))
    return f32(Q).astype(np.float64), f32(QD).astype(np.float64), f32(T).astype(np.float64)


# Contact-rich states of the stacking scene reach 60-96 contacts (pads on cube_1, cube_1 on cube_2, 16-20 cube-plane points),
# twice the row count the pick scene's floor was sized for; the fp32 build of the oracle differs from its fp64 build by up
# to 1.5e-5 rad/s on the golden trajectory (tools/make_stack_traj.py prints the spread).
TOL_DQD_FLOOR_STACK = 6e-5


def stack_substep_errors(q0, qd0, q_ref, qd_ref, q, qd, a_smooth_scale):
    eq = np.abs(q - q_ref).max(axis=1)
    d_ref = np.abs(qd_ref - qd0).max(axis=1)
    ev = np.abs(qd - qd_ref).max(axis=1)
    bound = TOL_DQD_REL * np.maximum(d_ref, a_smooth_scale) + TOL_DQD_FLOOR_STACK
    return eq * TOL_Q_ABS / (TOL_Q_ABS + 0.01 * bound), ev / bound


def stack_cluster_states(g, idx=(30, 36, 60, 80)):
    """Golden grasp / carry samples with the four other cubes lined up against cube_1 (0.0398 m pitch: 0.2 mm of overlap between
    neighbours): every body touches another one, so the cooperative solve runs on all 39 dofs (its two-rows-per-lane path)."""
    idx = list(idx)
    Q, QD, QA, T = (g[k][idx].copy() for k in ("q", "qd", "qacc", "act"))
    for r in range(len(idx)):
        c1 = Q[r, 9:12].copy()
        for k, off in zip(range(1, 5), (0.0398, -0.0398, 0.0796, -0.0796)):
            Q[r, 9 + 7 * k: 9 + 7 * k + 7] = [c1[0] + off, c1[1], 0.0199 if r < len(idx) - 1 else c1[2], 1, 0, 0, 0]
        QD[r, 9:] = 0
        QA[r] = 0
    return tuple(f32(x).astype(np.float64) for x in (Q, QD, QA, T))


def stack_pile_states(m, orc, n=12, steps=90, every=15, seed=7):
    """Tumbling piles: the five cubes are dropped from 5-25 cm above one spot with random orientations and spins and the oracle
    rolls the scene forward; the sampled states contain edge-edge and vertex-face cube-cube contacts, cubes resting on two others,
    and cubes bouncing off the plane on one corner.  Returns float32-rounded (Q, QD, QA, T) stacked over the sample times."""
    rng = np.random.default_rng(seed)
    nu, K = m.nu, m.n_obj
    Q, QD, QA = orc.new_state(n)
    Q[:, :nu] = m.home_q
    for k in range(K):
        o = nu + 7 * k
        Q[:, o:o + 2] = np.array([0.55, 0.0]) + rng.uniform(-0.03, 0.03, (n, 2))
        Q[:, o + 2] = m.plane_z + 0.05 + 0.05 * k + rng.uniform(0, 0.01, n)
        qq = rng.normal(size=(n, 4))
        Q[:, o + 3:o + 7] = qq / np.linalg.norm(qq, axis=1, keepdims=True)
        QD[:, nu + 6 * k:nu + 6 * k + 3] = rng.normal(0, 0.2, (n, 3))
        QD[:, nu + 6 * k + 3:nu + 6 * k + 6] = rng.normal(0, 3.0, (n, 3))
    T = np.tile(m.home_q, (n, 1))
    out = []
    for s in range(steps):
        if s % every == every - 1:
            out.append((f32(Q).astype(np.float64), f32(QD).astype(np.float64), f32(QA).astype(np.float64), T.copy()))
        orc.step(Q, QD, QA, T, iterations=100, tol=1e-10)
    return tuple(np.concatenate([o[i] for o in out]) for i in range(4))


def stack_clutter_states(m, gg_model, rng, n):
    """Cubes dropped around a random pad of the gripper at random poses (overlapping the fingers, each other and the plane, up to
    ~75 contacts) with random velocities: every kind of body pair of the stacking scene in one state."""
    nu, K = m.nu, m.n_obj
    Q = np.zeros((n, nu + 7 * K)); QD = np.zeros((n, nu + 6 * K))
    Q[:, :nu] = m.home_q + rng.uniform(-0.6, 0.6, (n, nu))
    Q[:, m.n_arm:nu] = rng.uniform(0, 0.04, (n, m.n_fing))
    Q[:, :nu] = np.clip(Q[:, :nu], np.array(list(gg_model.lim_lo)[:nu]), np.array(list(gg_model.lim_hi)[:nu]))
    pads = [g for g in m.geoms if g.group == 0]
    for k in range(n):
        xp, xm = m.fk(np.concatenate([Q[k, :nu], m.default_cube_q()]))
        g = pads[rng.integers(len(pads))]
        tip = xp[g.body] + xm[g.body] @ np.asarray(g.pos)
        for c in range(K):
            o = nu + 7 * c
            Q[k, o:o + 3] = tip + rng.uniform(-0.05, 0.05, 3) if (c < 2 or rng.random() < 0.5) else \
                [rng.uniform(0.3, 0.7), rng.uniform(-0.3, 0.3), m.plane_z + 0.0199]
            Q[k, o + 2] = max(Q[k, o + 2], m.plane_z + 0.0199)
            qq = rng.normal(size=4)
            Q[k, o + 3:o + 7] = qq / np.linalg.norm(qq)
            QD[k, nu + 6 * c:nu + 6 * c + 6] = rng.normal(0, 0.3, 6)
    QD[:, :nu] = rng.normal(0, 0.5, (n, nu))
    T = m.home_q + rng.uniform(-0.5, 0.5, (n, nu))
    T[:, m.n_arm:] = rng.uniform(-0.02, 0.04, (n, m.n_fing))
    return tuple(f32(x).astype(np.float64) for x in (Q, QD, T))


def pick_clutter_states(m, gg_model, rng, n):
    """CubePick scene with the cube dropped around a random pad of the gripper at a random pose (overlapping the fingers and / or the
    plane) and random velocities: the coupled path of the step kernel (warp-per-env solver teams) on states the scripted grasp never
    visits."""
    nu = m.nu
    Q = np.zeros((n, nu + 7)); QD = np.zeros((n, nu + 6))
    Q[:, :nu] = m.home_q + rng.uniform(-0.6, 0.6, (n, nu))
    if m.name == "franka":
        Q[:, m.n_arm:nu] = rng.uniform(0, 0.04, (n, m.n_fing))
    Q[:, :nu] = np.clip(Q[:, :nu], np.array(list(gg_model.lim_lo)[:nu]), np.array(list(gg_model.lim_hi)[:nu]))
    pads = [g for g in m.geoms if g.group == 0]
    for k in range(n):
        xp, xm = m.fk(m.full_q(Q[k, :nu]))
        g = pads[rng.integers(len(pads))]
        tip = xp[g.body] + xm[g.body] @ np.asarray(g.pos)
        Q[k, nu:nu + 3] = tip + rng.uniform(-0.04, 0.04, 3)
        Q[k, nu + 2] = max(Q[k, nu + 2], m.plane_z + 0.0199)
        qq = rng.normal(size=4)
        Q[k, nu + 3:nu + 7] = qq / np.linalg.norm(qq)
    QD[:, :nu] = rng.normal(0, 0.5, (n, nu))
    QD[:, nu:] = rng.normal(0, 0.3, (n, 6))
    T = m.home_q + rng.uniform(-0.5, 0.5, (n, nu))
    return tuple(f32(x).astype(np.float64) for x in (Q, QD, T))


# ---- numpy restatement of the device random streams (Philox4x32-10, Random123) -------------------------------------------------------
# Vectorised over envs so that the full-size tests can reproduce every env's per-step reset draw; pinned to the oracle's generator
# (tests/test_gpu_fullsize_parity.py::test_philox_numpy_matches_oracle_generator), which is pinned to the Random123 known answers.
def philox4x32_10_np(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32).astype(np.uint64) for c in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0), np.uint64(k1)
    M0, M1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ k0
        n1 = p1 & MASK
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ k1
        n3 = p0 & MASK
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + np.uint64(0x9E3779B9)) & MASK
        k1 = (k1 + np.uint64(0xBB67AE85)) & MASK
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def u01_np(x):
    """uint32 -> float32 in [0, 1): top 24 bits * 2^-24 (gg_math.cuh: u01)."""
    return (np.asarray(x, dtype=np.uint32) >> np.uint32(8)).astype(np.float32) * np.float32(5.9604644775390625e-8)
