"""CPU tier of the batched IK: the oracle's dense damped-least-squares restatement of
`robot.inverse_kinematics` against (i) an independent numpy Gauss-Newton IK, (ii) forward kinematics of its own answer, and
the device header gg_ik.cuh compiled for the host (tests/hostsim) against the oracle."""
import ctypes as C

import numpy as np
import pytest

from gym_genesis_b200 import cabi
from gym_genesis_b200.model import franka_cube_pick_model, get_model, quat_to_mat
from helpers import f32, ik_numpy
from oracle.binding import Oracle

TOL_IK_Q = 2e-4      # rad: fp32 device header vs fp64 oracle, same number of updates (the DLS map is a contraction)


def reachable_targets(m, orc, rng, n, spread=0.25):
    """Targets = FK of random configurations near home, starts = home + smaller noise: always reachable."""
    qt = np.tile(m.home_q, (n, 1)) + rng.uniform(-spread, spread, (n, m.nu))
    qt = np.clip(qt, np.asarray(m.flat()["lim_lo"])[:m.nu] + 1e-3, np.asarray(m.flat()["lim_hi"])[:m.nu] - 1e-3)
    tp, tq = np.zeros((n, 3)), np.zeros((n, 4))
    for i in range(n):
        xp, xq = orc.fk(m.full_q(qt[i]))
        tp[i], tq[i] = xp[m.eef_body], xq[m.eef_body]
    q0 = np.tile(m.home_q, (n, 1)) + rng.uniform(-0.05, 0.05, (n, m.nu))
    return f32(q0).astype(np.float64), f32(tp).astype(np.float64), f32(tq).astype(np.float64)


@pytest.mark.parametrize("robot", ["franka", "so101"])
def test_oracle_ik_reaches_target(robot):
    m = get_model(robot)
    orc = Oracle(m, np.float64)
    rng = np.random.RandomState(3)
    q0, tp, tq = reachable_targets(m, orc, rng, 16, spread=0.15 if robot == "so101" else 0.25)
    kw = dict(max_iters=200, pos_tol=1e-6, rot_tol=1e-5) if robot == "so101" else dict()
    # the 5-dof SO-101 cannot reach an arbitrary 6-dof pose, but these targets ARE poses of the arm, so the error still goes to 0
    q, err, iters = orc.ik(q0, tp, tq, **kw)
    ptol, rtol = (1e-6, 1e-5) if robot == "so101" else (5e-4, 5e-3)
    assert (err[:, 0] < ptol).all() and (err[:, 1] < rtol).all(), err.max(0)
    for i in range(4):
        xp, xq = orc.fk(m.full_q(q[i]))
        assert np.abs(xp[m.eef_body] - tp[i]).max() < 2 * ptol
    assert (q[:, m.n_arm:] == q0[:, m.n_arm:]).all()           # gripper dofs untouched
    assert iters.max() <= (200 if robot == "so101" else 20)


def test_oracle_ik_agrees_with_numpy_gauss_newton():
    """The expert's grasp pose (hand pointing down above a cube): both solvers must land on the same solution branch (the one
    closest to home), the oracle with Genesis' damping, numpy with 1e-6.  The 7-dof arm is redundant, so the two damping
    values may stop at slightly different points of the 1-dim self-motion manifold: joints agree to 0.3 rad (elbow-up branch near home; the step clamp also shapes the path), poses to 1e-7."""
    m = franka_cube_pick_model()
    orc = Oracle(m, np.float64)
    R = quat_to_mat([0, 1, 0, 0])
    for xy in [(0.5, 0.0), (0.65, 0.2), (0.70, -0.15)]:
        tp = np.array([[xy[0], xy[1], 0.02 + 0.115]])
        q, err, it = orc.ik(m.home_q[None], tp, np.array([[0.0, 1.0, 0.0, 0.0]]), max_iters=100, pos_tol=1e-9, rot_tol=1e-9)
        qn = ik_numpy(m, tp[0], R, m.home_q.copy(), iters=100)
        assert err[0, 0] < 1e-8 and err[0, 1] < 1e-8
        assert np.abs(q[0, :7] - qn[:7]).max() < 0.3, (q[0], qn)
        xa, _ = orc.fk(m.full_q(q[0])); xb, _ = orc.fk(m.full_q(qn))
        assert np.abs(xa[m.eef_body] - tp[0]).max() < 1e-7 and np.abs(xb[m.eef_body] - tp[0]).max() < 1e-7


def test_oracle_ik_masks_step_clamp_and_limits():
    m = franka_cube_pick_model()
    orc = Oracle(m, np.float64)
    # position-only target: orientation left free
    tp = np.array([[0.55, 0.1, 0.3]])
    q, err, _ = orc.ik(m.home_q[None], tp, None)
    assert err[0, 0] < 5e-4 and err[0, 1] == 0.0
    # one update with a far target: every dof moves by at most max_step
    q1, _, it = orc.ik(m.home_q[None], np.array([[0.3, 0.5, 0.8]]), np.array([[0.0, 1.0, 0.0, 0.0]]), max_iters=1, max_step=0.05)
    assert it[0] == 1 and np.abs(q1[0] - m.home_q).max() <= 0.05 + 1e-12
    # unreachable target (2 m away): stays inside the joint limits, reports a large error, uses all iterations
    q2, err2, it2 = orc.ik(m.home_q[None], np.array([[2.0, 0.0, 0.5]]), None)
    lo, hi = np.asarray(m.flat()["lim_lo"])[:9], np.asarray(m.flat()["lim_hi"])[:9]
    assert (q2[0] >= lo - 1e-6).all() and (q2[0] <= hi + 1e-6).all() and err2[0, 0] > 0.5 and it2[0] == 20
    # z-only mask
    q3, err3, _ = orc.ik(m.home_q[None], np.array([[9.0, 9.0, 0.35]]), None, mask6=[0, 0, 1, 0, 0, 0])
    xp, _ = orc.fk(m.full_q(q3[0]))
    assert abs(xp[m.eef_body][2] - 0.35) < 5e-4


def _cfg(**kw):
    k = cabi.GgIkConfig(20, 0.01, 5e-4, 5e-3, 0.5, 1, (C.c_float * 3)(1, 1, 1), (C.c_float * 3)(1, 1, 1))
    for a, b in kw.items():
        setattr(k, a, b)
    return k


@pytest.mark.parametrize("robot", ["franka", "so101"])
def test_device_header_ik_matches_oracle(robot, hostsim):
    m = get_model(robot)
    orc = Oracle(m, np.float64)
    gm = cabi.compile_gg(m)
    rng = np.random.RandomState(11)
    n = 64
    q0, tp, tq = reachable_targets(m, orc, rng, n, spread=0.15 if robot == "so101" else 0.25)
    hs = hostsim
    for iters in (1, 3, 20):
        # fixed number of updates (tolerances 0): the fp32 Gram-matrix formulation must track the fp64 dense one
        k = _cfg(max_iters=iters, pos_tol=0.0, rot_tol=0.0)
        qo = np.zeros((n, m.nu), np.float32); err = np.zeros((n, 2), np.float32); it = np.zeros(n, np.int32)
        a, b, c = f32(q0), f32(tp), f32(tq)
        hs.hs_ik(C.byref(gm), C.byref(k), n, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p),
                 qo.ctypes.data_as(C.c_void_p), err.ctypes.data_as(C.c_void_p), it.ctypes.data_as(C.c_void_p))
        qr, er, ir = orc.ik(q0, tp, tq, max_iters=iters, pos_tol=0.0, rot_tol=0.0)
        assert (it == iters).all() and (ir == iters).all()
        assert np.abs(qo - qr).max() < TOL_IK_Q, (iters, np.abs(qo - qr).max())
        assert np.abs(err - er).max() < 1e-4
    # default tolerances: same iteration counts except where the error sits within fp32 rounding of the threshold
    k = _cfg()
    hs.hs_ik(C.byref(gm), C.byref(k), n, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p),
             qo.ctypes.data_as(C.c_void_p), err.ctypes.data_as(C.c_void_p), it.ctypes.data_as(C.c_void_p))
    qr, er, ir = orc.ik(q0, tp, tq)
    same = it == ir
    assert same.mean() > 0.9
    assert np.abs(qo[same] - qr[same]).max() < TOL_IK_Q
    if robot == "franka":
        assert (err[:, 0] < 5e-4).all() and (err[:, 1] < 5e-3).all()


def bad_start_problem(m, orc, rng, n):
    """Reachable targets, start configurations drawn uniformly inside the joint limits: the first attempt fails for ~60 % of the rows
    within 20 updates, so the re-sampling path (max_samples) decides the result."""
    q0, tp, tq = reachable_targets(m, orc, rng, n, spread=0.25)
    lo, hi = np.asarray(m.flat()["lim_lo"])[:m.n_arm], np.asarray(m.flat()["lim_hi"])[:m.n_arm]
    q0[:, :m.n_arm] = rng.uniform(lo, hi, (n, m.n_arm))
    return f32(q0).astype(np.float64), tp, tq


def test_ik_resampling_first_success_in_sample_order(hostsim):
    """max_samples: unsolved rows restart from Philox draws (row, sample) inside the joint limits; result = first success in sample order.
    Oracle (fp64, sequential) against the device header (fp32, sequential form of ik_resample_kernel): same draws, same choice."""
    m = franka_cube_pick_model()
    orc = Oracle(m, np.float64)
    gm = cabi.compile_gg(m)
    n = 64
    q0, tp, tq = bad_start_problem(m, orc, np.random.RandomState(3), n)
    q1, e1, i1 = orc.ik(q0, tp, tq)
    ok1 = (e1[:, 0] < 5e-4) & (e1[:, 1] < 5e-3)
    assert 0.2 < ok1.mean() < 0.7                                  # the first attempt alone fails for many rows
    seed = (9 << 32) | 7
    qr, er, ir = orc.ik(q0, tp, tq, max_samples=16, seed=seed)
    okr = (er[:, 0] < 5e-4) & (er[:, 1] < 5e-3)
    assert okr.mean() > 0.95 and (ir[~ok1] > 20).all()              # re-sampling solves them, at the cost of further attempts
    assert np.array_equal(qr[ok1], q1[ok1])                         # rows solved by attempt 0 are untouched
    assert np.array_equal(orc.ik(q0, tp, tq, max_samples=16, seed=seed)[0], qr)          # deterministic in (seed, row, sample)
    assert not np.array_equal(orc.ik(q0, tp, tq, max_samples=16, seed=seed + 1)[0][~ok1], qr[~ok1])
    k = _cfg(max_samples=16, seed_lo=seed & 0xFFFFFFFF, seed_hi=seed >> 32)
    qo = np.zeros((n, m.nu), np.float32); err = np.zeros((n, 2), np.float32); it = np.zeros(n, np.int32)
    a, b, c = f32(q0), f32(tp), f32(tq)
    hostsim.hs_ik(C.byref(gm), C.byref(k), n, a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), c.ctypes.data_as(C.c_void_p),
                  qo.ctypes.data_as(C.c_void_p), err.ctypes.data_as(C.c_void_p), it.ctypes.data_as(C.c_void_p))
    same = it == ir                                                 # same attempts, same number of updates: the same sample was chosen
    assert same.mean() > 0.85, same.mean()
    assert np.abs(qo[same] - qr[same]).max() < 5 * TOL_IK_Q         # random starts amplify fp32 rounding a little more than near-home starts
    assert ((err[:, 0] < 5e-4) & (err[:, 1] < 5e-3)).mean() > 0.95
