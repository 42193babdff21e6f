"""The switchable Genesis-compatibility risks, oracle + device pipeline:
  * GENESIS_COMPAT_PYRAMID_MU_SCALING: regulariser of a pyramidal contact row, R = reg_scale (1 - imp) / imp (w_a + w_b) with
    reg_scale = 2 mu^2 (1 + mu^2) (on, MuJoCo lineage) or 1 (off).  Observable: resting penetration depth, closed form
    m g = n_rows D K imp depth with D = 1 / R -- 16 mm for the SO-101 task's friction 5 when on, 13 micrometres when off.
  * GENESIS_COMPAT_PANDA_WRIST_FORCE: actuator clamp of Panda joints 5-7 (87 N m as set by the stacking tasks, 12 N m as in the public
    MJCF).  Observable: the clamped PD torque and the joint acceleration it produces.
CPU tier on the host build of the device header; GPU tier through the C ABI."""
import ctypes as C

import numpy as np
import pytest

from gym_genesis_b200 import cabi
from gym_genesis_b200 import model as mdl
from gym_genesis_b200.model_so101 import ISLAND_TOP_Z, so101_cube_pick_model
from helpers import TOL_Q_ABS, _p, f32, hostsim_step, random_states, substep_errors


def _oracle(m):
    from oracle.binding import Oracle
    return Oracle(m, np.float64)


def _scale(o, Q, QD, T):
    return np.array([np.abs(o.dynamics(Q[k], QD[k], T[k])["a_smooth"]).max() * 0.01 for k in range(Q.shape[0])])


def _resting_depth(m):
    """closed form of the soft contact: 4 corners x 4 pyramidal rows, D = 1 / R, K = 1 / (dmax^2 tc^2 dr^2), imp = dmax at depth > width"""
    w = 1.0 / m.bodies[m.cube_body].mass
    imp, K = mdl.GENESIS_COMPAT_DMAX, 1.0 / (mdl.GENESIS_COMPAT_DMAX ** 2 * mdl.GENESIS_COMPAT_TIMECONST ** 2)
    R = m.reg_scale() * (1 - imp) / imp * w
    return m.bodies[m.cube_body].mass * mdl.GENESIS_COMPAT_GRAVITY / (16.0 / R * K * imp)


@pytest.mark.parametrize("scaling", [True, False])
def test_pyramid_mu_scaling_resting_depth_oracle(scaling):
    m = so101_cube_pick_model(success_height=ISLAND_TOP_Z + 0.1, pyramid_mu_scaling=scaling)
    assert m.reg_scale() == (np.float32(2 * 25.0 * 26.0) if scaling else 1.0)
    o = _oracle(m)
    Q, QD, QA = o.new_state(1)
    o.reset(Q, QD, QA, np.array([[-0.3, 0.0]]))
    T = m.home_q[None].copy()
    for _ in range(400):
        o.step(Q, QD, QA, T, iterations=50, tol=1e-10)
    depth = m.plane_z + 0.02 - Q[0, m.nu + 2]
    ref = _resting_depth(m)
    if scaling:
        assert 0.014 < ref < 0.018 and abs(depth - ref) < 0.05 * ref, (depth, ref)
    else:
        # imp is on the lower branch of the impedance curve (depth < width = 1 mm): the closed form with imp = dmax is a lower bound
        assert depth < 1e-4 and depth > 0.5 * ref, (depth, ref)


@pytest.mark.parametrize("scaling", [True, False])
def test_pyramid_mu_scaling_device_pipeline_matches_oracle(hostsim, scaling):
    m = so101_cube_pick_model(success_height=ISLAND_TOP_Z + 0.1, pyramid_mu_scaling=scaling)
    o, gg = _oracle(m), cabi.compile_gg(m)
    assert gg.reg_scale == np.float32(m.reg_scale())
    rng = np.random.default_rng(12)
    Q, QD, T = random_states(m, rng, 48, cube_on_ground=True)
    QA = np.zeros_like(QD)
    qo, qdo, qao = Q.copy(), QD.copy(), QA.copy()
    info, _ = o.step(qo, qdo, qao, T, iterations=100, tol=1e-12)
    q, qd, qa, hinfo, _ = hostsim_step(hostsim, gg, Q, QD, QA, T)
    assert (hinfo[:, 0] == info[:, 0]).all()
    eq, ev = substep_errors(Q, QD, qo, qdo, q, qd, _scale(o, Q, QD, T))
    assert eq.max() < TOL_Q_ABS and ev.max() < 1.0, (eq.max(), ev.max())


def test_franka_mu_scaling_off_changes_the_contact_stiffness_by_four(franka_model):
    on, off = franka_model, mdl.franka_cube_pick_model(pyramid_mu_scaling=False)
    assert on.reg_scale() == 4.0 and off.reg_scale() == 1.0           # mu = 1: 2 mu^2 (1 + mu^2) = 4
    assert abs(_resting_depth(on) / _resting_depth(off) - 4.0) < 1e-12


@pytest.mark.parametrize("wrist", [87.0, 12.0])
def test_wrist_force_range_switch(hostsim, wrist):
    m = mdl.franka_cube_pick_model(wrist_force=wrist)
    t = m.dof_table()
    assert list(t["frc_hi"][:9]) == [87.0] * 4 + [wrist] * 3 + [100.0] * 2 and list(t["frc_lo"][4:7]) == [-wrist] * 3
    o, gg = _oracle(m), cabi.compile_gg(m)
    rng = np.random.default_rng(5)
    Q, QD, T = random_states(m, rng, 48, cube_on_ground=True)
    T[:, 4:7] = Q[:, 4:7] + rng.choice([-0.3, 0.3], (48, 3))          # kp 2000 x 0.3 rad = 600 N m of PD demand: clamped either way
    d = o.dynamics(Q[0], QD[0], T[0])
    assert np.all(np.abs(d["tau"][4:7] + m.dof_table()["damping"][4:7] * QD[0, 4:7]) <= wrist + 1e-9)
    assert np.any(np.abs(np.abs(d["tau"][4:7] + QD[0, 4:7]) - wrist) < 1e-6)          # ... and sits ON the clamp
    QA = np.zeros_like(QD)
    qo, qdo, qao = Q.copy(), QD.copy(), QA.copy()
    o.step(qo, qdo, qao, T, iterations=100, tol=1e-12)
    q, qd, qa, hinfo, _ = hostsim_step(hostsim, gg, Q, QD, QA, T)
    eq, ev = substep_errors(Q, QD, qo, qdo, q, qd, _scale(o, Q, QD, T))
    assert eq.max() < TOL_Q_ABS and ev.max() < 1.0, (eq.max(), ev.max())
    if wrist == 12.0:                                                  # the observable: wrist accelerations drop with the clamp
        o87 = _oracle(mdl.franka_cube_pick_model(wrist_force=87.0))
        a12 = np.abs(o.dynamics(Q[0], QD[0], T[0])["a_smooth"][4:7]).max()
        a87 = np.abs(o87.dynamics(Q[0], QD[0], T[0])["a_smooth"][4:7]).max()
        assert a12 < 0.5 * a87


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["wrist12", "so101_mu_off"])
def test_switches_on_the_gpu(which):
    import torch
    from gym_genesis_b200.sim import BatchedSim
    m = mdl.franka_cube_pick_model(wrist_force=12.0) if which == "wrist12" else \
        so101_cube_pick_model(success_height=ISLAND_TOP_Z + 0.1, pyramid_mu_scaling=False)
    o = _oracle(m)
    rng = np.random.default_rng(31)
    n = 256
    Q, QD, T = random_states(m, rng, n, cube_on_ground=True)
    if which == "wrist12":
        T[:, 4:7] = Q[:, 4:7] + rng.choice([-0.3, 0.3], (n, 3))
    QA = np.zeros_like(QD)
    sim = BatchedSim(m, n)
    sim.set_state(Q, QD, QA)
    sim.step(torch.as_tensor(f32(T)))
    q, qd, _ = (t.cpu().numpy().astype(np.float64) for t in sim.get_state())
    qo, qdo, qao = Q.copy(), QD.copy(), QA.copy()
    o.step(qo, qdo, qao, T, iterations=100, tol=1e-12)
    eq, ev = substep_errors(Q, QD, qo, qdo, q, qd, _scale(o, Q, QD, T))
    assert eq.max() < TOL_Q_ABS and ev.max() < 1.0, (eq.max(), ev.max())
    sim.close()
