"""GPU tier, BASELINE.json sizes: oracle parity of the CubePick step at 4096 envs (configs[1]) and 16384 envs with masked partial
resets every step (configs[2]), and of the SO-101 articulation at 16384 envs (configs[3]).

Every step of the run:
  * ALL envs: reset decision (caller mask | auto-reset of finished envs | per-step Bernoulli draw of the device Philox stream),
    step / episode counters, terminated / truncated / reward are reproduced BIT-EXACTLY on the host from the documented rules
    (reference semantics: franka/cube_pick.py:86-135, env.py:63-65; partial resets: DESIGN.md section 1) -- the Philox draws come
    from a numpy restatement pinned to the oracle's generator (itself pinned to the Random123 known-answer vectors);
  * a subset of >= 512 envs -- the last (partial) block, every env that reset in this step, random others -- is advanced by the
    fp64 oracle from the GPU's own pre-step state and compared with the single-substep bound of tests/helpers.py;
  * ALL envs: observations against the oracle's observation function evaluated on the GPU's post-step state.
"""
import numpy as np
import pytest
import torch

from helpers import TOL_Q_ABS, f32, philox4x32_10_np, pick_clutter_states, substep_errors, u01_np

pytestmark = pytest.mark.gpu

RESET_TAG = 0x52535421      # counter word 2 of the per-step reset draw (gg_kernels.cu: reset_decision)


def _np(t):
    return t.detach().cpu().numpy()


def _spawn_state(m, orc, seed, gids, episodes):
    """Reset state of the envs `gids` entering episode `episodes` + 1: home pose, zero velocities / warm start, cube at the device
    Philox spawn (seed, global env id, episode) with the task's spawn height and quaternion."""
    k = len(gids)
    Q, QD, QA = orc.new_state(k)
    Q[:, :m.nu] = m.home_q
    for r, (g, ep) in enumerate(zip(gids, episodes)):
        Q[r, m.nu:m.nu + 2] = orc.spawn_xy(seed, int(g), int(ep)).astype(np.float64)
    Q[:, m.nu + 2] = np.float32(m.spawn_z)
    Q[:, m.nu + 3:m.nu + 7] = f32(m.spawn_quat)
    return Q, QD, QA


def _run_fullsize(Sim, m, orc, n, steps, seed, reset_prob, auto_reset, time_limit, amp, subset, first_env_id=0, host_mask_every=0):
    rng = np.random.default_rng(seed + 1000)
    sim = Sim(m, n, seed=seed, reset_prob=reset_prob, auto_reset=auto_reset, time_limit=time_limit, first_env_id=first_env_id)
    sim.reset(None, None)
    nu = m.nu
    step_index = 0
    assert (_np(sim.episode) == 1).all() and (_np(sim.step_count) == 0).all()
    done_prev = np.zeros(n, bool)                       # library-owned `done` is cleared by the full reset
    gid = (first_env_id + np.arange(n)).astype(np.uint32)
    success_z = np.float32(m.success_z)
    home32 = f32(m.home_q)
    worst_q = worst_v = 0.0
    n_reset_total = n_checked = 0
    for k in range(steps):
        q0, qd0, qa0 = (_np(t).astype(np.float64) for t in sim.get_state())
        sc0, ep0 = _np(sim.step_count).copy(), _np(sim.episode).copy()
        act = home32 + rng.uniform(-amp, amp, (n, nu)).astype(np.float32)
        if m.name == "franka":
            act[:, 7:] = rng.uniform(-0.02, 0.04, (n, 2)).astype(np.float32)
        mask = None
        if host_mask_every and k % host_mask_every == host_mask_every - 1:
            mask = rng.random(n) < 0.01
        sim.step(torch.as_tensor(act), reset_mask=None if mask is None else torch.as_tensor(mask))
        # ---- host restatement of the reset decision for ALL envs
        do_reset = np.zeros(n, bool) if mask is None else mask.copy()
        if auto_reset:
            do_reset |= done_prev
        if reset_prob > 0:
            o = philox4x32_10_np(gid, np.full(n, step_index, np.uint32), np.full(n, RESET_TAG, np.uint32), np.zeros(n, np.uint32),
                                 np.uint32(seed & 0xFFFFFFFF), np.uint32(seed >> 32))
            do_reset |= u01_np(o[0]) < np.float32(reset_prob)
        step_index += 1
        q1, qd1, qa1 = (_np(t).astype(np.float64) for t in sim.get_state())
        sc1, ep1 = _np(sim.step_count), _np(sim.episode)
        assert np.array_equal(ep1, ep0 + do_reset), f"step {k}: episode counters differ on {(ep1 != ep0 + do_reset).sum()} envs"
        assert np.array_equal(sc1, np.where(do_reset, 0, sc0 + 1)), f"step {k}: step counters"
        z32 = q1[:, nu + 2].astype(np.float32)
        succ = z32 > success_z
        trunc = (sc1 >= m.horizon) if time_limit else np.zeros(n, bool)
        assert np.array_equal(_np(sim.terminated).astype(bool), succ), f"step {k}: terminated"
        assert np.array_equal(_np(sim.truncated).astype(bool), trunc), f"step {k}: truncated"
        assert np.array_equal(_np(sim.reward), succ.astype(np.float32)), f"step {k}: reward"
        done_prev = succ | trunc
        # ---- observations of ALL envs from the GPU's own post-step state
        a, e, r, t = orc.obs(q1)
        assert np.abs(_np(sim.obs_agent) - a).max() < 2e-6 and np.abs(_np(sim.obs_env) - e).max() < 2e-6, f"step {k}: observations"
        assert np.array_equal(t, succ)
        # ---- oracle substep on the subset
        ridx = np.nonzero(do_reset)[0]
        n_reset_total += len(ridx)
        tail = np.arange(max(0, n - 32), n)
        others = rng.choice(n, size=subset, replace=False)
        idx = np.unique(np.concatenate([tail, ridx[:128], others]))[: max(subset, 512) + 160]
        is_r = do_reset[idx]
        Q, QD, QA = q0[idx].copy(), qd0[idx].copy(), qa0[idx].copy()
        T = act[idx].astype(np.float64)
        if is_r.any():
            Qs, QDs, QAs = _spawn_state(m, orc, seed, gid[idx][is_r], ep0[idx][is_r])
            Q[is_r], QD[is_r], QA[is_r] = Qs, QDs, QAs
            T[is_r] = home32.astype(np.float64)
        Qin, QDin = Q.copy(), QD.copy()
        stepped = ~is_r if m.settle_steps == 0 else np.ones(len(idx), bool)
        if stepped.any():
            qs, qds, qas = Q[stepped].copy(), QD[stepped].copy(), QA[stepped].copy()
            orc.step(qs, qds, qas, T[stepped], iterations=100, tol=1e-12)
            Q[stepped], QD[stepped] = qs, qds
        if (~stepped).any():                              # reset without a settle step: the reset state itself, bit for bit
            assert np.array_equal(f32(q1[idx][~stepped]), f32(Q[~stepped])) and np.abs(qd1[idx][~stepped]).max() == 0.0
        sel = np.nonzero(stepped)[0]
        scale = np.array([np.abs(orc.dynamics(Qin[j], QDin[j], T[j])["a_smooth"]).max() * 0.01 for j in sel])
        eq, ev = substep_errors(Qin[sel], QDin[sel], Q[sel], QD[sel], q1[idx][sel], qd1[idx][sel], scale)
        worst_q, worst_v = max(worst_q, eq.max()), max(worst_v, ev.max())
        assert eq.max() < TOL_Q_ABS and ev.max() < 1.0, (k, eq.max(), ev.max(), int(idx[sel][np.argmax(ev)]))
        n_checked += len(sel)
    st = _np(sim.stats())
    assert st[5] == 0                                   # no non-finite env was re-spawned
    sim.close()
    return dict(resets=n_reset_total, checked=n_checked, worst_q=worst_q, worst_v=worst_v)


@pytest.fixture(scope="module")
def Sim():
    from gym_genesis_b200.sim import BatchedSim
    return BatchedSim


def test_philox_numpy_matches_oracle_generator():
    from oracle.binding import philox
    rng = np.random.default_rng(0)
    c = rng.integers(0, 2 ** 32, (64, 4), dtype=np.uint64).astype(np.uint32)
    key = rng.integers(0, 2 ** 32, 2, dtype=np.uint64).astype(np.uint32)
    o = philox4x32_10_np(c[:, 0], c[:, 1], c[:, 2], c[:, 3], key[0], key[1])
    for r in range(64):
        assert np.array_equal(np.array([o[i][r] for i in range(4)], np.uint32), philox(c[r], key))


def test_config2_4096_envs_oracle_parity(Sim, franka_model, oracle64):
    """BASELINE configs[1]: CubePick-v0 Franka, 4096 envs on one GPU, 12 steps, no resets after the first."""
    r = _run_fullsize(Sim, franka_model, oracle64, 4096, 12, seed=11, reset_prob=0.0, auto_reset=False, time_limit=False, amp=0.3, subset=512)
    assert r["checked"] >= 12 * 512


def test_config3_16384_envs_masked_resets_every_step(Sim, franka_model, oracle64):
    """BASELINE configs[2]: 16384 envs, randomized cube poses, Bernoulli(1/200) partial resets EVERY step + auto-reset + time limit,
    a caller-supplied mask on top every third step; global env ids offset as on rank 1 of a sharded run."""
    r = _run_fullsize(Sim, franka_model, oracle64, 16384, 12, seed=(5 << 32) | 1234, reset_prob=1.0 / 200, auto_reset=True, time_limit=True,
                      amp=0.3, subset=512, first_env_id=16384, host_mask_every=3)
    assert r["resets"] > 12 * 50 and r["checked"] >= 12 * 512


def test_config3_partial_tail_block(Sim, franka_model, oracle64):
    """Same rules with an env count that is not a multiple of the block's 32 envs (the last block is partial)."""
    r = _run_fullsize(Sim, franka_model, oracle64, 16384 - 19, 6, seed=77, reset_prob=1.0 / 200, auto_reset=True, time_limit=False,
                      amp=0.3, subset=512)
    assert r["resets"] > 6 * 40


def test_config4_so101_16384_envs_masked_resets(Sim, so101_model, so101_oracle64):
    """BASELINE configs[3]: SO-101 articulation (5 hinges + jaw), 16384 envs, partial resets every step (no settle step in this task:
    a reset env holds its reset state for the step, so101/cube_pick.py:81)."""
    r = _run_fullsize(Sim, so101_model, so101_oracle64, 16384, 10, seed=9, reset_prob=1.0 / 200, auto_reset=True, time_limit=False,
                      amp=0.1, subset=512)
    assert r["resets"] > 10 * 40 and r["checked"] >= 10 * 512


# ---- the thread-per-env kernel (launch_step picks it above one wave of role blocks = 128 envs per SM, unless more than a few per cent of the
#      envs are coupled; GG_STEP_KERNEL=mono forces it at any size) ---------------------------------------------------------------------
@pytest.fixture
def mono_kernel(monkeypatch):
    monkeypatch.setenv("GG_STEP_KERNEL", "mono")
    yield
    monkeypatch.delenv("GG_STEP_KERNEL", raising=False)


def test_mono_kernel_oracle_parity_with_resets(mono_kernel, Sim, franka_model, oracle64):
    r = _run_fullsize(Sim, franka_model, oracle64, 4096 + 7, 8, seed=21, reset_prob=1.0 / 100, auto_reset=True, time_limit=True, amp=0.3,
                      subset=512, host_mask_every=3)
    assert r["resets"] > 8 * 20 and r["checked"] >= 8 * 512


def test_mono_kernel_coupled_envs_through_the_grid_wide_list(mono_kernel, Sim, franka_model, oracle64):
    """Cube at random poses around the fingers (pads overlapping cube and / or plane): the coupled envs of the thread-per-env kernel go
    through the grid-wide list and the warp-per-env drain (gg_drain.cuh); same bound as the role kernel's clutter test."""
    from helpers import pick_clutter_states
    n = 1024
    sim = Sim(franka_model, n)
    Q, QD, T = pick_clutter_states(franka_model, sim.gg_model, np.random.default_rng(57), n)
    QA = np.zeros_like(QD)
    sim.set_state(Q, QD, QA)
    sim.step(torch.as_tensor(f32(T)))
    q, qd, _ = (_np(t).astype(np.float64) for t in sim.get_state())
    qo, qdo, qao = Q.copy(), QD.copy(), QA.copy()
    info, _ = oracle64.step(qo, qdo, qao, T, iterations=300, tol=1e-12)
    assert (info[:, 0] > 4).sum() > 200
    scale = np.array([np.abs(oracle64.dynamics(Q[k], QD[k], T[k])["a_smooth"]).max() * 0.01 for k in range(n)])
    eq, ev = substep_errors(Q, QD, qo, qdo, q, qd, scale)
    assert eq.max() < TOL_Q_ABS and ev.max() < 1.0, (eq.max(), ev.max(), int(np.argmax(ev)))
    a, e, r, t = oracle64.obs(q)
    assert np.abs(_np(sim.obs_agent) - a).max() < 2e-6 and np.abs(_np(sim.obs_env) - e).max() < 2e-6
    assert np.array_equal(_np(sim.terminated).astype(bool), t)
    st = _np(sim.stats())
    assert st[0] == n and st[5] == 0 and st[7] / st[0] > 4          # every env stepped exactly once; contact-rich on average
    sim.close()


def test_kernel_dispatch_by_env_count(Sim, franka_model):
    """131072 envs on one GPU take the thread-per-env kernel automatically; the result obeys the same invariants."""
    n = 131072
    sim = Sim(franka_model, n, seed=3, reset_prob=1.0 / 200, auto_reset=True)
    sim.reset(None, None)
    g = torch.Generator(device="cuda").manual_seed(0)
    home = torch.as_tensor(franka_model.home_q, dtype=torch.float32, device="cuda")
    for k in range(10):
        act = home + 0.6 * (torch.rand((n, 9), generator=g, device="cuda") - 0.5)
        act[:, 7:] = -0.02 + 0.06 * torch.rand((n, 2), generator=g, device="cuda")
        sim.step(act)
    q, qd, _ = sim.get_state()
    assert torch.isfinite(q).all() and torch.isfinite(qd).all()
    assert (q[:, 12:16].norm(dim=1) - 1).abs().max() < 1e-5 and q[:, 11].min() > 0.0195 and q[:, 11].max() < 0.021
    st = _np(sim.stats())
    assert st[5] == 0 and st[0] > 0.98 * n * 10 * (1 - 1 / 200.0)
    sim.close()


def test_adaptive_kernel_choice_follows_the_coupled_fraction(Sim, franka_model):
    """Above one wave of role blocks the thread-per-env kernel runs the free-space steps; once the (lagged) fraction of coupled envs passes
    4 % launch_step switches to the role kernel, whose block-local solver teams are faster on contact-rich steps.  Both kernels write
    bit-identical states, so the switch must be invisible: same final state as a run forced onto the role kernel."""
    import ctypes as C
    import os
    from gym_genesis_b200 import cabi
    n_sm = torch.cuda.get_device_properties(0).multi_processor_count
    tile = 512
    n = ((128 * n_sm) // tile + 2) * tile                         # just above one wave of role blocks
    rng = np.random.default_rng(11)
    Q, QD, T = pick_clutter_states(franka_model, cabi.compile_gg(franka_model), rng, tile)
    Qn, QDn, Tn = (np.tile(x, (n // tile, 1)) for x in (Q, QD, T))
    home = np.tile(f32(franka_model.home_q), (n, 1))

    def run(forced):
        if forced:
            os.environ["GG_STEP_KERNEL"] = forced
        try:
            sim = Sim(franka_model, n, seed=4)
        finally:
            os.environ.pop("GG_STEP_KERNEL", None)
        sim.lib.gg_debug_dispatch.argtypes = [C.c_void_p, C.c_void_p]
        probe = (C.c_float * 3)()
        trace = []
        sim.reset(None, None)
        a_free, a_cl = torch.as_tensor(home, device="cuda"), torch.as_tensor(f32(Tn), device="cuda")
        for k in range(48):                                       # free space: nothing coupled
            sim.step(a_free)
        torch.cuda.synchronize()
        sim.lib.gg_debug_dispatch(sim._h, probe); trace.append(tuple(probe))
        for rep in range(3):                                      # contact rich: every env starts with the cube between / around the pads
            sim.set_state(f32(Qn), f32(QDn), np.zeros_like(f32(QDn)))
            for k in range(34):
                sim.step(a_cl)
                torch.cuda.synchronize()
        sim.lib.gg_debug_dispatch(sim._h, probe); trace.append(tuple(probe))
        out = [t.clone() for t in sim.get_state()]
        sim.close()
        return trace, out

    trace, out = run(None)
    assert trace[0][0] == 1.0 and trace[0][1] == 0.0 and trace[0][2] == 0.0          # thread-per-env kernel, nothing coupled seen
    assert trace[1][1] == 1.0 and trace[1][2] > 0.04, trace                           # switched to the role kernel
    _, ref = run("roles")
    for x, y in zip(out, ref):
        assert torch.equal(x, y)
