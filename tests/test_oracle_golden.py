"""Pins the oracle and the model tables against every known answer available for this path
(the reference has no tests; known answers are the published Franka ready pose,
the reference home pose, the SO-101 XML-derived FK vectors and the Random123 Philox KAT) and against
physics identities that do not depend on Genesis."""
import json
import os

import numpy as np
import pytest

from gym_genesis_b200 import model as mdl
from oracle.binding import philox

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "fk_vectors.json")))


def test_philox_known_answers():
    for kat in GOLD["philox4x32_10"]:
        out = philox(kat["ctr"], kat["key"])
        assert [f"{int(x):08x}" for x in out] == kat["out"]


def test_franka_ready_pose(franka_model, oracle64):
    g = GOLD["franka_ready"]
    q = franka_model.full_q(np.array(g["q"] + [0.04, 0.04]))
    for fk in (lambda q: franka_model.fk(q)[0], lambda q: oracle64.fk(q)[0]):
        xp = fk(q)
        assert np.allclose(xp[8], g["hand_pos"], atol=2e-6)
    xp, xm = franka_model.fk(q)
    assert np.allclose(xp[8] + xm[8] @ [0, 0, 0.1034], g["tcp_pos"], atol=2e-6)
    assert np.allclose(np.diag(xm[8]), g["hand_R_diag"], atol=1e-6)


def test_franka_home_pose(franka_model, oracle64):
    g = GOLD["franka_home"]
    q = franka_model.full_q(np.array(g["q"] + [0.04, 0.04]))
    xp, xq = oracle64.fk(q)
    assert np.allclose(xp[8], g["hand_pos"], atol=2e-6)
    ref = np.array(g["hand_quat_wxyz_upto_sign"])
    assert min(np.abs(xq[8] - ref).max(), np.abs(xq[8] + ref).max()) < 1e-4


def test_oracle_mass_matrix_matches_numpy_model(franka_model, oracle64):
    rng = np.random.default_rng(0)
    for _ in range(5):
        qr = mdl.FRANKA_HOME + rng.uniform(-0.5, 0.5, 9)
        qr[7:] = rng.uniform(0, 0.04, 2)
        cq = mdl.quat_normalize(rng.normal(size=4))
        q = franka_model.full_q(qr, np.array([0.6, 0.1, 0.3, *cq]))
        M = oracle64.dynamics(q, np.zeros(15), mdl.FRANKA_HOME)["M"]
        Mn = franka_model.mass_matrix(q)
        assert np.abs(M - Mn).max() < 2e-6            # constants are fp32-rounded in the oracle
        assert np.allclose(M, M.T, atol=1e-12)
        assert np.linalg.eigvalsh(M).min() > 0


def test_oracle_gravity_term_is_potential_gradient(franka_model, oracle64):
    """bias(q, 0) = dV/dq with V = sum m g z_com, finite differences on the numpy model."""
    m = franka_model
    rng = np.random.default_rng(1)
    qr = mdl.FRANKA_HOME + rng.uniform(-0.5, 0.5, 9)
    q = m.full_q(qr)

    def V(qv):
        xp, xm = m.fk(qv)
        return sum(b.mass * m.gravity * (xp[i] + xm[i] @ np.asarray(b.com))[2] for i, b in enumerate(m.bodies))

    bias = oracle64.dynamics(q, np.zeros(15), mdl.FRANKA_HOME)["bias"]
    for d in range(9):
        dq = np.zeros_like(q); dq[d] = 1e-6
        g = (V(q + dq) - V(q - dq)) / 2e-6
        assert abs(bias[d] - g) < 5e-5, (d, bias[d], g)
    assert abs(bias[11] - m.bodies[m.cube_body].mass * m.gravity) < 1e-6


def test_oracle_coriolis_matches_christoffel(franka_model, oracle64):
    """bias(q, qd) - bias(q, 0) = sum_jk (dM_ij/dq_k - 1/2 dM_jk/dq_i) qd_j qd_k (robot dofs; numpy M, central differences)."""
    m = franka_model
    rng = np.random.default_rng(2)
    qr = mdl.FRANKA_HOME + rng.uniform(-0.5, 0.5, 9)
    q = m.full_q(qr)
    qd = np.zeros(15); qd[:9] = rng.normal(size=9)
    h = 1e-5
    dM = []
    for k in range(9):
        dq = np.zeros_like(q); dq[k] = h
        dM.append((m.mass_matrix(q + dq)[:9, :9] - m.mass_matrix(q - dq)[:9, :9]) / (2 * h))
    dM = np.array(dM)            # dM[k, i, j] = dM_ij/dq_k
    v = qd[:9]
    c = np.einsum("kij,j,k->i", dM, v, v) - 0.5 * np.einsum("ijk,j,k->i", dM, v, v)
    b1 = oracle64.dynamics(q, qd, mdl.FRANKA_HOME)["bias"][:9]
    b0 = oracle64.dynamics(q, np.zeros(15), mdl.FRANKA_HOME)["bias"][:9]
    assert np.abs((b1 - b0) - c).max() < 2e-4 * max(1.0, np.abs(c).max())


def test_oracle_invweight_matches_host_compiler(franka_model, oracle64):
    bw, dw = oracle64.invweight(franka_model.full_q(np.zeros(9)))
    bw2, dw2 = franka_model.invweight0()
    moving = bw2 > 1e-12
    assert np.allclose(bw[moving], bw2[moving], rtol=1e-5)
    assert np.allclose(dw, dw2, rtol=1e-5)
    assert abs(bw2[franka_model.cube_body] - 1 / 0.0128) < 1e-9


def test_free_fall_closed_form(franka_model, oracle64):
    o = oracle64
    Q, QD, QA = o.new_state(1)
    o.reset(Q, QD, QA, np.array([[0.6, 0.0]]))
    Q[0, 11] = 0.5
    tgt = mdl.FRANKA_HOME[None].copy()
    o.step(Q, QD, QA, tgt)
    assert abs(QD[0, 11] + 9.81 * 0.01) < 1e-6
    assert abs(Q[0, 11] - (0.5 - 9.81 * 1e-4)) < 1e-7


def test_resting_contact_closed_form(franka_model, oracle64):
    """Steady state on the plane: 16 pyramidal rows share m g  =>  depth = m g / (16 D K imp(depth))."""
    o = oracle64
    Q, QD, QA = o.new_state(1)
    o.reset(Q, QD, QA, np.array([[0.6, 0.0]]))
    tgt = mdl.FRANKA_HOME[None].copy()
    for _ in range(300):
        info, _ = o.step(Q, QD, QA, tgt)
    assert info[0, 0] == 4 and abs(QD[0, 11]) < 1e-9
    depth = 0.02 - Q[0, 11]
    J, aref, D = o.rows(Q[0], QD[0])
    rows = J[:, 11] > 0.5          # the 16 contact rows all have a unit z entry
    assert rows.sum() == 16
    # force balance: sum_rows D * aref = m g
    assert abs((D[rows] * aref[rows]).sum() - 0.0128 * 9.81) < 1e-7
    assert 5e-5 < depth < 3e-4
    # friction: a small tangential push inside the pyramid does not move the cube
    assert np.abs(QD[0, 9:11]).max() < 1e-9


def test_newton_reaches_kkt(franka_model, oracle64):
    """At the oracle's solution the gradient of the convex cost vanishes: Mt (a - a_s) = J^T f, f = -D min(0, J a - aref)."""
    from helpers import scripted_grasp_states
    states, reward = scripted_grasp_states(franka_model, oracle64, stride=20)
    assert reward == 1.0            # the scripted grasp lifts the cube above z = 0.1 in the oracle
    o = oracle64
    for Q, QD, QA, T in states:
        d = o.dynamics(Q[0], QD[0], T[0])
        J, aref, D = o.rows(Q[0], QD[0])
        q, qd, qa = Q.copy(), QD.copy(), QA.copy()
        o.step(q, qd, qa, T, iterations=200, tol=1e-14)
        a = qa[0]
        jar = J @ a - aref
        f = -D * np.minimum(jar, 0)
        g = d["Mt"] @ (a - d["a_smooth"]) - J.T @ f
        assert np.abs(g).max() < 1e-7 * max(1.0, np.abs(d["Mt"] @ d["a_smooth"]).max())


def test_so101_fk_known_answers():
    """SO-101 tree against vectors computed from the reference's own MJCF (tests/golden/make_so101_fk.py)."""
    from gym_genesis_b200.model import BoxGeom, SceneModel, _cube_body, mat_to_quat
    from gym_genesis_b200.model_so101 import SO101_STACK_HOME, so101_bodies
    from oracle.binding import Oracle
    b = so101_bodies()
    b.append(_cube_body())
    sm = SceneModel("so101raw", b, [BoxGeom(7, (0.02,) * 3, group=1)], eef_body=5, cube_body=7, home_q=np.zeros(6), n_arm=5, n_fing=1)
    orc = Oracle(sm, np.float64)
    assert abs(sum(x.mass for x in b[:7]) - GOLD["so101_zero"]["total_mass"]) < 1e-6
    for key, q in (("so101_zero", np.zeros(6)), ("so101_stack_home", SO101_STACK_HOME)):
        g = GOLD[key]
        xp, xq = orc.fk(sm.full_q(q))
        assert np.allclose(xp[5], g["gripper_pos"], atol=2e-6)
        ref = np.array(g["gripper_quat"])
        assert min(np.abs(xq[5] - ref).max(), np.abs(xq[5] + ref).max()) < 2e-6
        if "jaw_pos" in g:
            assert np.allclose(xp[6], g["jaw_pos"], atol=2e-6)
        xpn, xmn = sm.fk(sm.full_q(q))
        assert np.allclose(xpn, xp, atol=1e-6)


def test_so101_oracle_matches_numpy_mass_matrix(so101_model, so101_oracle64):
    rng = np.random.default_rng(5)
    m = so101_model
    for _ in range(3):
        q = m.full_q(m.home_q + rng.uniform(-0.3, 0.3, 6), np.array([-0.3, 0.0, 0.9, 1.0, 0, 0, 0]))
        M = so101_oracle64.dynamics(q, np.zeros(12), m.home_q)["M"]
        Mn = m.mass_matrix(q)
        assert np.abs(M - Mn).max() < 1e-6 * max(1.0, np.abs(Mn).max()) and np.linalg.eigvalsh(M).min() > 0
