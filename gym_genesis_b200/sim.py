"""Torch-facing host wrapper of the C-ABI simulator (`include/gg_b200.h`).

PyTorch is plumbing here: it owns the device memory (state, observations, flags) and supplies the
CUDA stream; every computation of the step path happens inside `libgg_b200.so`.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import cabi
from .model import SceneModel


class GgError(RuntimeError):
    pass


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _storage_uses(t) -> int:
    return torch._C._storage_Use_Count(t.untyped_storage()._cdata)


class _SlabRing:
    """Hands out per-step tensor sets cut from K-step slabs and recycles a slab once nobody references its storage any more."""
    MAX_SLABS = 4

    def __init__(self, make_bases, split):
        self.make_bases, self.split = make_bases, split
        self.slabs = []             # [base tensors, storage use counts with no views alive]
        self.cur = -1
        self.pool = []

    def next(self):
        if not self.pool:
            n = len(self.slabs)
            for i in range(1, n + 1):               # the other slabs first: the current one still backs the latest results
                j = (self.cur + i) % n
                bases, idle = self.slabs[j]
                if all(_storage_uses(b) <= c for b, c in zip(bases, idle)):
                    self.cur = j
                    break
            else:
                bases = self.make_bases()
                slab = [bases, [_storage_uses(b) for b in bases]]
                if n == 0:
                    # a steady loop alternates between two slabs: make both on the first call (the caller's warm-up), so that no
                    # allocation (pinned slabs: milliseconds of page locking) lands K steps into a loop somebody is timing
                    spare = self.make_bases()
                    self.slabs.append([spare, [_storage_uses(b) for b in spare]])
                    n = 1
                if n < self.MAX_SLABS:
                    self.slabs.append(slab)
                    self.cur = n
                else:                               # the caller keeps old results alive: they own those slabs now
                    self.cur = (self.cur + 1) % n
                    self.slabs[self.cur] = slab
            self.pool = self.split(self.slabs[self.cur][0])
        return self.pool.pop()


class BatchedSim:
    """N independent worlds (arm + gripper + free cube + plane) stepped by one CUDA kernel per env.step.

    Replaces, for this path, the Genesis `Scene`/`RigidSolver` objects the reference task holds
    (`gym_genesis/tasks/franka/cube_pick.py:37-68`).
    """

    def __init__(self, scene_model: SceneModel, n_envs: int, device="cuda", solver_iterations: int = 10,
                 solver_tolerance: float = 1e-5, auto_reset: bool = False, time_limit: bool = False, reset_prob: float = 0.0,
                 seed: int = 0, first_env_id: int = 0, fresh_outputs: bool = False):
        if not torch.cuda.is_available():
            raise GgError("gym_genesis_b200 needs a CUDA device: the env.step path has no CPU fallback")
        self.lib = cabi.load_library()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError(f"device must be a CUDA device, got {device}")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.scene_model = scene_model
        self.gg_model = cabi.compile_gg(scene_model)
        self.n_envs = int(n_envs)
        self.nu = scene_model.nu
        self.fresh_outputs = bool(fresh_outputs)
        self.env_dim = 11
        self.nq, self.nv = self.nu + 7, self.nu + 6
        self.agent_dim = 7 + scene_model.n_fing
        h = C.c_void_p()
        self._check(self.lib.gg_create(C.byref(self.gg_model), self.n_envs, self.device.index, C.byref(h)))
        self._h = h
        rows, stride = self.lib.gg_state_rows(h), self.lib.gg_state_stride(h)
        dev = self.device
        # struct-of-arrays state: [field component][env], env fastest (coalesced per warp)
        self.state = torch.zeros((rows, stride), dtype=torch.float32, device=dev)
        self.step_count = torch.zeros(self.n_envs, dtype=torch.int32, device=dev)
        self.episode = torch.zeros(self.n_envs, dtype=torch.int32, device=dev)
        self._check(self.lib.gg_bind_state(h, _ptr(self.state), _ptr(self.step_count), _ptr(self.episode)))
        self.obs_agent = torch.zeros((self.n_envs, self.agent_dim), dtype=torch.float32, device=dev)
        self.obs_env = torch.zeros((self.n_envs, 11), dtype=torch.float32, device=dev)
        self.reward = torch.zeros(self.n_envs, dtype=torch.float32, device=dev)
        # torch.bool is one byte holding 0/1: the kernel writes the flags in place, no conversion kernels on the step path
        self.terminated = torch.zeros(self.n_envs, dtype=torch.bool, device=dev)
        self.truncated = torch.zeros(self.n_envs, dtype=torch.bool, device=dev)
        self._stats = torch.zeros(8, dtype=torch.float64, device=dev)
        self._out_ptrs = tuple(t.data_ptr() for t in (self.obs_agent, self.obs_env, self.reward, self.terminated, self.truncated))
        self.cfg = cabi.GgConfig(solver_iterations, solver_tolerance, int(auto_reset), int(time_limit), float(reset_prob),
                                 int(seed) & (2 ** 64 - 1), int(first_env_id))
        self._check(self.lib.gg_configure(h, C.byref(self.cfg)))

    # ------------------------------------------------------------------------------------------
    def _check(self, rc):
        if rc != 0:
            raise GgError(self.lib.gg_last_error().decode())

    def _stream(self):
        # raw handle of torch's CURRENT stream on this device (the public torch.cuda.current_stream() builds a Stream object: 4 us per
        # step on the host-I/O path, where the whole Python side of a step is ~15 us)
        try:
            return torch._C._cuda_getCurrentRawStream(self.device.index)
        except AttributeError:                              # pragma: no cover - older / newer torch without the private accessor
            return torch.cuda.current_stream(self.device).cuda_stream

    def configure(self, **kw):
        for k, v in kw.items():
            if not hasattr(self.cfg, k):
                raise AttributeError(k)
            setattr(self.cfg, k, v)
        self._check(self.lib.gg_configure(self._h, C.byref(self.cfg)))

    def close(self):
        if getattr(self, "_h", None) is not None:
            self.lib.gg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------
    def _as_action(self, action) -> torch.Tensor:
        if not isinstance(action, torch.Tensor):
            action = torch.as_tensor(np.asarray(action), dtype=torch.float32)
        if action.shape != (self.n_envs, self.nu):
            raise ValueError(f"action must have shape ({self.n_envs}, {self.nu}), got {tuple(action.shape)}")
        if action.device != self.device or action.dtype != torch.float32 or not action.is_contiguous():
            action = action.to(device=self.device, dtype=torch.float32, non_blocking=True).contiguous()
        return action

    def _as_mask(self, mask) -> Optional[torch.Tensor]:
        if mask is None:
            return None
        if not isinstance(mask, torch.Tensor):
            mask = torch.as_tensor(np.asarray(mask))
        if mask.shape != (self.n_envs,):
            raise ValueError(f"mask must have shape ({self.n_envs},), got {tuple(mask.shape)}")
        return mask.to(device=self.device, dtype=torch.uint8).contiguous()

    # ---- fresh outputs ------------------------------------------------------------------------------------------------------------
    # fresh_outputs=True: every step / reset writes into output tensors nobody has seen before, so tensors returned earlier keep their
    # values -- the reference's getters return fresh tensors (franka/cube_pick.py:137-157) and its collectors append them step after
    # step (examples/franka/pick_cube_state.py:86-100).  Creating five tensors per step costs ~25 us of host time (allocator + Python
    # tensor objects), as much as the step kernel itself; so the sets are made in SLABS: one allocation per output kind for K steps,
    # split into K views by a single unbind() call, and a step just takes the next set.  Slabs are recycled (`_SlabRing`): once every
    # view of a slab has been dropped by the caller -- the storage use count is back to its baseline -- the slab is split again, so a
    # steady loop allocates nothing (pinned allocations are driver calls that take milliseconds when several ranks share a host).
    # A caller that keeps results simply makes the ring allocate new slabs.  The default (False) reuses one set of buffers:
    # zero allocation, CUDA-graph safe.
    _POOL_BYTES = 64 << 20

    def _pool_size(self):
        per_step = self.n_envs * (self.agent_dim + self.env_dim + 1) * 4 + 2 * self.n_envs
        return int(max(1, min(32, self._POOL_BYTES // per_step)))

    # The three rings are built on first use; a step only pops the next prepared set (views + their raw pointers, computed from the slab's
    # base pointer so that splitting a slab costs one unbind() per tensor kind and no per-view data_ptr() call).
    def _make_out(self):
        n, dev, K = self.n_envs, self.device, self._pool_size()
        return (torch.empty((K, n, self.agent_dim), dtype=torch.float32, device=dev), torch.empty((K, n, self.env_dim), dtype=torch.float32, device=dev),
                torch.empty((K, n), dtype=torch.float32, device=dev), torch.empty((2, K, n), dtype=torch.bool, device=dev))

    def _split_out(self, bases):
        A, E, R = (b.unbind(0) for b in bases[:3])
        T, U = bases[3][0].unbind(0), bases[3][1].unbind(0)
        n, K = self.n_envs, len(A)
        pa, pe, pr, pt = (b.data_ptr() for b in bases)
        sa, se, sr, pu = 4 * n * self.agent_dim, 4 * n * self.env_dim, 4 * n, pt + K * n
        return [(A[k], E[k], R[k], T[k], U[k], (pa + k * sa, pe + k * se, pr + k * sr, pt + k * n, pu + k * n)) for k in range(K)]

    def _new_outputs(self):
        ring = self.__dict__.get("_out_ring")
        if ring is None:
            ring = self._out_ring = _SlabRing(self._make_out, self._split_out)
        pool = ring.pool
        self.obs_agent, self.obs_env, self.reward, self.terminated, self.truncated, self._out_ptrs = pool.pop() if pool else ring.next()

    def _make_host(self):
        n, K = self.n_envs, self._pool_size()
        return (torch.empty((K, n), dtype=torch.float32, pin_memory=True), torch.empty((2, K, n), dtype=torch.bool, pin_memory=True))

    def _split_host(self, bases):
        R, T, U = bases[0].unbind(0), bases[1][0].unbind(0), bases[1][1].unbind(0)
        n, K = self.n_envs, len(R)
        pr, pt = bases[0].data_ptr(), bases[1].data_ptr()
        pu = pt + K * n
        return [(R[k], T[k], U[k], pr + 4 * n * k, pt + n * k, pu + n * k) for k in range(K)]

    def _new_host_results(self):
        """Same slabs for the pinned host mirrors of step_host: the kernel writes each step's results into a pinned slot nobody has
        seen before, so the returned CPU tensors are fresh without a copy."""
        ring = self.__dict__.get("_host_ring")
        if ring is None:
            ring = self._host_ring = _SlabRing(self._make_host, self._split_host)
        pool = ring.pool
        r, t, u, pr, pt, pu = pool.pop() if pool else ring.next()
        hio, h = self._hio, self._host
        hio.reward, hio.terminated, hio.truncated = pr, pt, pu
        h["reward"], h["terminated"], h["truncated"] = r, t, u

    def _make_host_obs(self):
        n, K = self.n_envs, self._pool_size()
        return (torch.empty((K, n, self.agent_dim), dtype=torch.float32, pin_memory=True),
                torch.empty((K, n, self.env_dim), dtype=torch.float32, pin_memory=True))

    def _split_host_obs(self, bases):
        A, E = bases[0].unbind(0), bases[1].unbind(0)
        n = self.n_envs
        pa, pe = bases[0].data_ptr(), bases[1].data_ptr()
        return [(A[k], E[k], pa + 4 * n * self.agent_dim * k, pe + 4 * n * self.env_dim * k) for k in range(len(A))]

    def _new_host_obs(self):
        """Fresh pinned observation mirrors (host_obs=True callers): the kernel writes agent_pos / environment_state of the step
        straight into them, so a host-resident policy gets fresh CPU tensors without a copy."""
        ring = self.__dict__.get("_host_obs_ring")
        if ring is None:
            ring = self._host_obs_ring = _SlabRing(self._make_host_obs, self._split_host_obs)
        pool = ring.pool
        a, e, pa, pe = pool.pop() if pool else ring.next()
        self._host["obs_agent"], self._host["obs_env"] = a, e
        self._host_obs_ptrs = (pa, pe)

    def step(self, action, reset_mask=None):
        """One env.step (= one 0.01 s substep) for all envs; results land in the obs_agent / obs_env / reward / terminated /
        truncated tensors of this object (no host sync)."""
        a = self._as_action(action)
        m = self._as_mask(reset_mask)
        self._last_action = a          # PD targets persist (reference: control_dofs_position semantics)
        if self.fresh_outputs:
            self._new_outputs()
        self._check(self._step_fn()(self._h, a.data_ptr(), *self._out_ptrs, _ptr(m), self._stream()))

    def _step_fn(self):
        return self.lib.gg_step

    def _host_buffers(self):
        if getattr(self, "_host", None) is None:
            n = self.n_envs
            self._host = {
                "action": torch.empty((n, self.nu), dtype=torch.float32).pin_memory(),      # staging for unpinned callers
                "reward": torch.zeros(n, dtype=torch.float32).pin_memory(),
                "terminated": torch.zeros(n, dtype=torch.bool).pin_memory(),
                "truncated": torch.zeros(n, dtype=torch.bool).pin_memory(),
                "obs_agent": torch.zeros((n, self.agent_dim), dtype=torch.float32).pin_memory(),
                "obs_env": torch.zeros((n, self.env_dim), dtype=torch.float32).pin_memory(),
            }
            # device-resident copy of the action batch the last host-fed step consumed (written by the kernel): PD targets persist
            self._targets_dev = torch.zeros((n, self.nu), dtype=torch.float32, device=self.device)
            self._hio = cabi.GgHostIO(0, self._host["reward"].data_ptr(), self._host["terminated"].data_ptr(),
                                      self._host["truncated"].data_ptr(), 1)
            self._hio.action_dev = self._targets_dev.data_ptr()
            self._hio_ref = C.byref(self._hio)
            self._pinned_ok = {}
        return self._host

    def step_host(self, action, reset_mask=None, wait=True, host_obs=False):
        """env.step for a HOST-resident caller (the reference's numpy convention): `action` is a CPU float32 (n_envs, n_dof)
        array; a pinned torch tensor is read in place by the kernel over PCIe (zero-copy), anything else is first copied into
        a pinned staging buffer.  Returns pinned CPU tensors (reward, terminated, truncated) the kernel has written directly;
        with wait=True (default) they are valid on return.  Observations stay on the device (obs_agent / obs_env);
        host_obs=True additionally has the kernel write them into the pinned mirrors `host_obs_agent` / `host_obs_env`
        (what a host-resident policy reads every step).  The consumed action batch is also kept on the device (`_last_action`)."""
        h = self._host_buffers()
        if self.fresh_outputs:
            self._new_outputs()
            self._new_host_results()
        # tensors that passed the checks once are recognised by (object id, data pointer): `is_pinned()` is a driver query
        key = (id(action), action.data_ptr()) if isinstance(action, torch.Tensor) else None
        if key is not None and key in self._pinned_ok:
            a = action
        else:
            if tuple(getattr(action, "shape", ())) != (self.n_envs, self.nu):
                raise ValueError(f"action must have shape ({self.n_envs}, {self.nu}), got {tuple(getattr(action, 'shape', ()))}")
            if isinstance(action, torch.Tensor) and action.device.type == "cpu" and action.dtype == torch.float32 \
                    and action.is_contiguous() and action.is_pinned():
                a = action
                if len(self._pinned_ok) < 64:
                    self._pinned_ok[key] = action          # keeps the tensor alive, so the id cannot be recycled
            else:
                a = h["action"]
                a.copy_(torch.as_tensor(np.asarray(action) if not isinstance(action, torch.Tensor) else action, dtype=torch.float32))
        m = None if reset_mask is None else self._as_mask(reset_mask)
        self._last_action = self._targets_dev      # filled by the kernel from the batch it consumed
        self._hio.action = a.data_ptr()
        self._hio.wait = 1 if wait else 0
        if host_obs:
            if self.fresh_outputs:
                self._new_host_obs()
            elif getattr(self, "_host_obs_ptrs", None) is None:
                self._host_obs_ptrs = (h["obs_agent"].data_ptr(), h["obs_env"].data_ptr())
            self._hio.obs_agent, self._hio.obs_env = self._host_obs_ptrs
        else:
            self._hio.obs_agent = self._hio.obs_env = None
        rc = self._step_host_fn()(self._h, self._hio_ref, *self._out_ptrs, _ptr(m), self._stream())
        if rc != 0:
            self._check(rc)
        return h["reward"], h["terminated"], h["truncated"]

    @property
    def host_obs_agent(self):
        return self._host_buffers()["obs_agent"]

    @property
    def host_obs_env(self):
        return self._host_buffers()["obs_env"]

    def _step_host_fn(self):
        return self.lib.gg_step_host

    def reset(self, mask=None, cube_xy=None):
        m = self._as_mask(mask)
        xy = None
        if cube_xy is not None:
            xy = torch.as_tensor(np.asarray(cube_xy, dtype=np.float32) if not isinstance(cube_xy, torch.Tensor) else cube_xy)
            if xy.shape != (self.n_envs, 2):
                raise ValueError(f"cube_xy must have shape ({self.n_envs}, 2)")
            xy = xy.to(device=self.device, dtype=torch.float32).contiguous()
        self._last_action = None
        if self.fresh_outputs:
            self._new_outputs()
        self._check(self.lib.gg_reset(self._h, _ptr(m), _ptr(xy), _ptr(self.obs_agent), _ptr(self.obs_env), self._stream()))

    def observe(self):
        self._check(self.lib.gg_observe(self._h, _ptr(self.obs_agent), _ptr(self.obs_env), _ptr(self.reward), _ptr(self.terminated),
                                        self._stream()))

    def get_state(self):
        q = torch.empty((self.n_envs, self.nq), dtype=torch.float32, device=self.device)
        qd = torch.empty((self.n_envs, self.nv), dtype=torch.float32, device=self.device)
        qacc = torch.empty((self.n_envs, self.nv), dtype=torch.float32, device=self.device)
        self._check(self.lib.gg_get_state(self._h, _ptr(q), _ptr(qd), _ptr(qacc), self._stream()))
        return q, qd, qacc

    def set_state(self, q=None, qd=None, qacc=None):
        def prep(t, k):
            if t is None:
                return None
            t = torch.as_tensor(t, dtype=torch.float32).to(self.device).contiguous()
            if t.shape != (self.n_envs, k):
                raise ValueError(f"expected shape ({self.n_envs}, {k}), got {tuple(t.shape)}")
            return t
        q, qd, qacc = prep(q, self.nq), prep(qd, self.nv), prep(qacc, self.nv)
        self._check(self.lib.gg_set_state(self._h, _ptr(q), _ptr(qd), _ptr(qacc), self._stream()))

    def inverse_kinematics(self, pos=None, quat=None, init_qpos=None, envs_idx=None, respect_joint_limit=True,
                           max_solver_iters=20, damping=0.01, pos_tol=5e-4, rot_tol=5e-3, pos_mask=(True, True, True),
                           rot_mask=(True, True, True), max_step_size=0.5, return_error=False, return_iters=False, max_samples=1):
        """Batched IK of the observed end-effector link: one CUDA launch (`gg_ik`), no host sync.  Mirrors the keywords of
        Genesis' `RigidEntity.inverse_kinematics` as the reference's experts call it (examples/franka/pick_cube_state.py:46-51).
        pos (B, 3) / quat (B, 4) wxyz; init_qpos (B, n_dof) or None = the envs' current joint positions; envs_idx (B,) or None =
        envs 0..B-1.  Returns qpos (B, n_dof) [, err (B, 2) = |e_pos|, |e_rot|] [, iters (B,)]."""
        if pos is None and quat is None:
            raise ValueError("inverse_kinematics needs pos and/or quat")
        dev = self.device

        def prep(t, k):
            if t is None:
                return None
            t = torch.as_tensor(t, dtype=torch.float32).to(dev)
            if t.dim() == 1:
                t = t[None]
            if t.shape[-1] != k or t.dim() != 2:
                raise ValueError(f"expected shape (B, {k}), got {tuple(t.shape)}")
            return t.contiguous()
        pos, quat, init_qpos = prep(pos, 3), prep(quat, 4), prep(init_qpos, self.nu)
        idx = None
        if envs_idx is not None:
            idx = torch.as_tensor(envs_idx)
            if idx.numel() > 0 and not torch.cuda.is_current_stream_capturing():
                lo, hi = int(idx.min()), int(idx.max())          # host check (one sync): the kernel indexes the state with these
                if lo < 0 or hi >= self.n_envs:
                    raise ValueError(f"envs_idx out of range [0, {self.n_envs}): min {lo}, max {hi}")
            idx = idx.to(device=dev, dtype=torch.int32).reshape(-1).contiguous()
        rows = {t.shape[0] for t in (pos, quat, init_qpos) if t is not None}
        if idx is not None:
            rows.add(idx.shape[0])
        if len(rows) != 1:
            raise ValueError(f"pos / quat / init_qpos / envs_idx disagree on the batch size: {sorted(rows)}")
        B = rows.pop()
        if idx is None and init_qpos is None and B > self.n_envs:
            raise ValueError(f"batch of {B} rows exceeds num_envs={self.n_envs}")
        k = cabi.GgIkConfig()
        self.lib.gg_ik_defaults(C.byref(k))
        k.max_iters, k.damping, k.pos_tol, k.rot_tol, k.max_step = int(max_solver_iters), float(damping), float(pos_tol), float(rot_tol), float(max_step_size)
        k.respect_limits = int(bool(respect_joint_limit))
        k.max_samples = max(1, int(max_samples))
        k.seed_lo, k.seed_hi = int(self.cfg.seed) & 0xFFFFFFFF, (int(self.cfg.seed) >> 32) & 0xFFFFFFFF
        for i in range(3):
            k.pos_mask[i] = 1.0 if pos_mask[i] else 0.0
            k.rot_mask[i] = 1.0 if rot_mask[i] else 0.0
        q_out = torch.empty((B, self.nu), dtype=torch.float32, device=dev)
        err = torch.empty((B, 2), dtype=torch.float32, device=dev) if (return_error or k.max_samples > 1) else None
        iters = torch.empty(B, dtype=torch.int32, device=dev) if return_iters else None
        self._check(self.lib.gg_ik(self._h, B, _ptr(idx), _ptr(pos), _ptr(quat), _ptr(init_qpos), _ptr(q_out), _ptr(err), _ptr(iters),
                                   C.byref(k), self._stream()))
        out = (q_out,) + ((err,) if return_error else ()) + ((iters,) if return_iters else ())     # err may exist only for the re-sampling pass
        return out[0] if len(out) == 1 else out

    def stats(self, clear=False) -> torch.Tensor:
        """[env-steps, episodes finished, successes, return sum, length sum, non-finite resets, solver iterations, contacts]."""
        self._check(self.lib.gg_stats(self._h, _ptr(self._stats), int(clear), self._stream()))
        return self._stats

    @staticmethod
    def launch_count() -> int:
        return int(cabi.load_library().gg_launch_count())


class BatchedStackSim(BatchedSim):
    """N independent worlds of the stacking scene (arm + gripper + five free cubes + plane), `gg_stack_*` entry points.

    Replaces the Genesis `Scene` of `gym_genesis/tasks/franka/cube_stack.py:42-96`.  One env.step = four launches
    (thread-per-env dynamics, thread-per-cube register solves, warp-per-env contact solve, thread-per-env integration / observation); see
    `csrc/gg_stack.cu`.  `inverse_kinematics` runs the same `gg_ik` kernel as the pick scene on this sim's joint state.
    """

    def __init__(self, scene_model: SceneModel, n_envs: int, device="cuda", solver_iterations: int = 10,
                 solver_tolerance: float = 1e-5, auto_reset: bool = False, time_limit: bool = False, reset_prob: float = 0.0,
                 seed: int = 0, first_env_id: int = 0, fresh_outputs: bool = False):
        if not torch.cuda.is_available():
            raise GgError("gym_genesis_b200 needs a CUDA device: the env.step path has no CPU fallback")
        self.lib = cabi.load_library()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError(f"device must be a CUDA device, got {device}")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.scene_model = scene_model
        self.fresh_outputs = bool(fresh_outputs)
        self.gg_model = cabi.compile_gg(scene_model)
        self.gg_task = cabi.compile_stack_task(scene_model)
        self.n_envs = int(n_envs)
        self.nu = scene_model.nu
        self.n_obj = scene_model.n_obj
        self.nq, self.nv = self.nu + 7 * self.n_obj, self.nu + 6 * self.n_obj
        self.env_dim = cabi.GG_STACK_ENV_STATE_DIM
        h = C.c_void_p()
        self._check(self.lib.gg_stack_create(C.byref(self.gg_model), C.byref(self.gg_task), self.n_envs, self.device.index, C.byref(h)))
        self._h = h
        self.agent_dim = self.lib.gg_stack_agent_dim(h)       # 7 + n_fing (eef pose + gripper) or n_dof (joint positions)
        rows, stride = self.lib.gg_stack_state_rows(h), self.lib.gg_stack_state_stride(h)
        dev = self.device
        self.state = torch.zeros((rows, stride), dtype=torch.float32, device=dev)
        self.step_count = torch.zeros(self.n_envs, dtype=torch.int32, device=dev)
        self.episode = torch.zeros(self.n_envs, dtype=torch.int32, device=dev)
        self._check(self.lib.gg_stack_bind_state(h, _ptr(self.state), _ptr(self.step_count), _ptr(self.episode)))
        self.obs_agent = torch.zeros((self.n_envs, self.agent_dim), dtype=torch.float32, device=dev)
        self.obs_env = torch.zeros((self.n_envs, self.env_dim), dtype=torch.float32, device=dev)
        self.reward = torch.zeros(self.n_envs, dtype=torch.float32, device=dev)
        self.terminated = torch.zeros(self.n_envs, dtype=torch.bool, device=dev)
        self.truncated = torch.zeros(self.n_envs, dtype=torch.bool, device=dev)
        self._stats = torch.zeros(8, dtype=torch.float64, device=dev)
        self._out_ptrs = tuple(t.data_ptr() for t in (self.obs_agent, self.obs_env, self.reward, self.terminated, self.truncated))
        self.cfg = cabi.GgConfig(solver_iterations, solver_tolerance, int(auto_reset), int(time_limit), float(reset_prob),
                                 int(seed) & (2 ** 64 - 1), int(first_env_id))
        self._check(self.lib.gg_stack_configure(h, C.byref(self.cfg)))
        self._ik_h = None

    def configure(self, **kw):
        for k, v in kw.items():
            if not hasattr(self.cfg, k):
                raise AttributeError(k)
            setattr(self.cfg, k, v)
        self._check(self.lib.gg_stack_configure(self._h, C.byref(self.cfg)))

    def close(self):
        if getattr(self, "_ik_h", None) is not None:
            self.lib.gg_destroy(self._ik_h)
            self._ik_h = None
        if getattr(self, "_h", None) is not None:
            self.lib.gg_stack_destroy(self._h)
            self._h = None

    def _step_fn(self):
        return self.lib.gg_stack_step

    def _step_host_fn(self):
        return self.lib.gg_stack_step_host

    def reset(self, mask=None, obj_xy=None):
        m = self._as_mask(mask)
        xy = None
        if obj_xy is not None:
            xy = torch.as_tensor(np.asarray(obj_xy, dtype=np.float32) if not isinstance(obj_xy, torch.Tensor) else obj_xy)
            if xy.shape != (self.n_envs, self.n_obj, 2):
                raise ValueError(f"obj_xy must have shape ({self.n_envs}, {self.n_obj}, 2)")
            xy = xy.to(device=self.device, dtype=torch.float32).contiguous()
        self._last_action = None
        if self.fresh_outputs:
            self._new_outputs()
        self._check(self.lib.gg_stack_reset(self._h, _ptr(m), _ptr(xy), _ptr(self.obs_agent), _ptr(self.obs_env), self._stream()))

    def observe(self):
        self._check(self.lib.gg_stack_observe(self._h, _ptr(self.obs_agent), _ptr(self.obs_env), _ptr(self.reward),
                                              _ptr(self.terminated), self._stream()))

    def get_state(self):
        q = torch.empty((self.n_envs, self.nq), dtype=torch.float32, device=self.device)
        qd = torch.empty((self.n_envs, self.nv), dtype=torch.float32, device=self.device)
        qacc = torch.empty((self.n_envs, self.nv), dtype=torch.float32, device=self.device)
        self._check(self.lib.gg_stack_get_state(self._h, _ptr(q), _ptr(qd), _ptr(qacc), self._stream()))
        return q, qd, qacc

    def set_state(self, q=None, qd=None, qacc=None):
        def prep(t, k):
            if t is None:
                return None
            t = torch.as_tensor(t, dtype=torch.float32).to(self.device).contiguous()
            if t.shape != (self.n_envs, k):
                raise ValueError(f"expected shape ({self.n_envs}, {k}), got {tuple(t.shape)}")
            return t
        q, qd, qacc = prep(q, self.nq), prep(qd, self.nv), prep(qacc, self.nv)
        self._check(self.lib.gg_stack_set_state(self._h, _ptr(q), _ptr(qd), _ptr(qacc), self._stream()))

    def step_info(self):
        """(contacts, Newton iterations) of the last step per env, int32 device tensors."""
        nc = torch.empty(self.n_envs, dtype=torch.int32, device=self.device)
        it = torch.empty(self.n_envs, dtype=torch.int32, device=self.device)
        self._check(self.lib.gg_stack_info(self._h, _ptr(nc), _ptr(it), self._stream()))
        return nc, it

    def inverse_kinematics(self, *args, **kw):
        # the IK kernel reads joint positions from state rows [0, n_dof) with the env index fastest: the stacking layout
        # starts with the same rows, so a CubePick handle bound to this sim's buffers serves it unchanged
        if self._ik_h is None:
            h = C.c_void_p()
            self._check(self.lib.gg_create(C.byref(self.gg_model), self.n_envs, self.device.index, C.byref(h)))
            assert self.lib.gg_state_stride(h) == self.state.shape[1]
            self._check(self.lib.gg_bind_state(h, _ptr(self.state), _ptr(self.step_count), _ptr(self.episode)))
            self._ik_h = h
        stack_h, self._h = self._h, self._ik_h
        try:
            return BatchedSim.inverse_kinematics(self, *args, **kw)
        finally:
            self._h = stack_h

    def stats(self, clear=False) -> torch.Tensor:
        self._check(self.lib.gg_stack_stats(self._h, _ptr(self._stats), int(clear), self._stream()))
        return self._stats
