"""Batched CubePick task classes on the B200 simulator.

Mirrors the task-class protocol of the reference (`gym_genesis/tasks/franka/cube_pick.py:21-181`):
`__init__(enable_pixels, observation_height, observation_width, num_envs, env_spacing,
camera_capture_mode, strip_environment_state)`, `reset() -> obs`, `seed(seed)`,
`step(action) -> (None, reward, None, obs)`, `compute_reward()`, `get_obs()`, attributes
`observation_space`, `action_space`, `scene`, robot / cube / eef handles.
"""
from __future__ import annotations

import random

import numpy as np
import torch

from ..gym_api import spaces
from ..model import get_model
from ..sim import BatchedSim

ENV_DIM = 11


class _Scene:
    """Stand-in for the Genesis scene handle: `scene.step()` advances physics holding the last targets
    (used by `GenesisEnv.push`, reference env.py:59-60)."""

    def __init__(self, task):
        self._task = task
        self.envs_offset = np.zeros((task.num_envs, 3))

    def step(self):
        self._task._physics_step(self._task._targets)


class _Link:
    def __init__(self, task):
        self._task = task

    def get_pos(self, envs_idx=None):
        p = self._task.sim.obs_agent[:, 0:3]
        return p.clone() if envs_idx is None else p[torch.as_tensor(envs_idx, device=p.device).long()]

    def get_quat(self, envs_idx=None):
        q = self._task.sim.obs_agent[:, 3:7]
        return q.clone() if envs_idx is None else q[torch.as_tensor(envs_idx, device=q.device).long()]


class _Robot:
    def __init__(self, task):
        self._task = task

    def get_link(self, name):
        return self._task.eef

    def get_dofs_position(self, envs_idx=None):
        q = self._task.sim.get_state()[0][:, : self._task.sim.nu]
        return q if envs_idx is None else q[torch.as_tensor(envs_idx, device=q.device).long()]

    get_qpos = get_dofs_position

    def get_dofs_velocity(self):
        return self._task.sim.get_state()[1][:, : self._task.sim.nu]

    def inverse_kinematics(self, link=None, pos=None, quat=None, init_qpos=None, respect_joint_limit=True, max_samples=50,
                           max_solver_iters=20, damping=0.01, pos_tol=5e-4, rot_tol=5e-3, pos_mask=(True, True, True),
                           rot_mask=(True, True, True), max_step_size=0.5, dofs_idx_local=None, return_error=False, envs_idx=None):
        """Genesis' `robot.inverse_kinematics(link=eef, pos=, quat=, envs_idx=)` as the reference experts call it
        (examples/franka/pick_cube_state.py:46-51): batched damped least squares on the GPU, returns qpos (B, n_dof).
        Only the end-effector link is supported.  `max_samples` (Genesis default 50): rows that miss the tolerances are re-started from
        random arm configurations drawn by the device Philox stream; the first success in sample order is returned (gg_ik.cuh)."""
        if link is not None and link is not self._task.eef:
            raise NotImplementedError("inverse_kinematics: only the end-effector link (get_link('hand')) is supported")
        if dofs_idx_local is not None:
            raise NotImplementedError("inverse_kinematics: dofs_idx_local is not supported")
        return self._task.sim.inverse_kinematics(pos=pos, quat=quat, init_qpos=init_qpos, envs_idx=envs_idx,
                                                 respect_joint_limit=respect_joint_limit, max_solver_iters=max_solver_iters,
                                                 damping=damping, pos_tol=pos_tol, rot_tol=rot_tol, pos_mask=pos_mask,
                                                 rot_mask=rot_mask, max_step_size=max_step_size, return_error=return_error,
                                                 max_samples=max_samples)

    def control_dofs_position(self, position, dofs_idx_local=None):
        idx = slice(None) if dofs_idx_local is None else torch.as_tensor(np.asarray(dofs_idx_local), device=self._task.sim.device)
        self._task._targets[:, idx] = torch.as_tensor(position, dtype=torch.float32, device=self._task.sim.device)


class _Cube:
    def __init__(self, task):
        self._task = task

    def get_pos(self):
        return self._task.sim.obs_env[:, 0:3].clone()

    def get_quat(self):
        return self._task.sim.obs_env[:, 3:7].clone()


class CubePickBatch:
    robot_name = "franka"

    def __init__(self, enable_pixels, observation_height, observation_width, num_envs, env_spacing, camera_capture_mode,
                 strip_environment_state, device="cuda", auto_reset=False, time_limit=False, solver_iterations=10,
                 solver_tolerance=1e-5, seed=0, first_env_id=0, reset_prob=0.0, success_height=None, alias_outputs=False):
        if enable_pixels:
            raise NotImplementedError("state-only path: rendering (enable_pixels=True) is out of scope (SURVEY.md section 5)")
        if num_envs <= 0:
            raise ValueError("num_envs must be positive for the batched task")
        self.enable_pixels = enable_pixels
        self.observation_height = observation_height
        self.observation_width = observation_width
        self.num_envs = num_envs
        self.camera_capture_mode = camera_capture_mode
        self.strip_environment_state = strip_environment_state
        self._random = np.random.RandomState()
        self.model = get_model(self.robot_name)
        if success_height is not None:
            self.model.success_z = float(success_height)
        elif self.model.success_z <= self.model.spawn_z:
            # the reference's literal threshold (so101/cube_pick.py:112: z > 0.1) with the cube spawned ABOVE it (stack placement: the island
            # top is at z = 0.7): reward = 1 and terminated = True on every step, and auto_reset would reset every step
            import warnings
            warnings.warn(f"{type(self).__name__}: success height {self.model.success_z} is below the cube's spawn height {self.model.spawn_z:.3f}; "
                          "every step reports success (the reference's literal threshold) -- pass success_height= to change it", stacklevel=3)
        # alias_outputs=False (default): every step / reset returns NEW tensors, as the reference's getters do (franka/cube_pick.py:137-157),
        # so `all_rewards.append(reward)` collectors and (obs, next_obs) replay buffers keep what they stored.  alias_outputs=True reuses
        # one set of output buffers (zero allocation, CUDA-graph safe): the tensors of step t are overwritten by step t + 1.
        self.alias_outputs = bool(alias_outputs)
        self.sim = BatchedSim(self.model, num_envs, device=device, solver_iterations=solver_iterations,
                              solver_tolerance=solver_tolerance, auto_reset=auto_reset, time_limit=time_limit,
                              reset_prob=reset_prob, seed=seed, first_env_id=first_env_id, fresh_outputs=not self.alias_outputs)
        self.agent_dim = self.sim.agent_dim
        self.observation_space = self._make_obs_space()
        self.action_space = spaces.Box(low=-1.0, high=1.0, shape=(self.sim.nu,), dtype=np.float32)
        self._home = torch.tensor(self.model.home_q, dtype=torch.float32, device=self.sim.device).repeat(num_envs, 1)
        self._targets = self._home.clone()
        self.scene = _Scene(self)
        self.eef = _Link(self)
        self.cube = _Cube(self)
        self.robot = _Robot(self)
        self.motors_dof = np.arange(self.model.n_arm)
        self.fingers_dof = np.arange(self.model.n_arm, self.model.nu)

    def _make_obs_space(self):
        return spaces.Dict({
            "agent_pos": spaces.Box(low=-np.inf, high=np.inf, shape=(self.agent_dim,), dtype=np.float32),
            "environment_state": spaces.Box(low=-np.inf, high=np.inf, shape=(ENV_DIM,), dtype=np.float32),
        })

    # ---- reference protocol ------------------------------------------------------------------
    def _spawn_xy(self):
        """Host RandomState stream of the reference: all x draws, then all y draws (cube_pick.py:90-91)."""
        B = self.num_envs
        lo, hi = self.model.spawn_lo, self.model.spawn_hi
        x = self._random.uniform(lo[0], hi[0], size=(B,))
        y = self._random.uniform(lo[1], hi[1], size=(B,))
        return np.stack([x, y], axis=1).astype(np.float32)

    def reset(self, mask=None):
        """Full reset (mask=None) reproduces cube_pick.py:86-112: spawn, home pose, hold-home targets, one settle step.
        A boolean mask resets only those envs (new capability); their spawn comes from the device Philox stream."""
        if mask is None:
            self.sim.reset(None, self._spawn_xy())
            self._targets = self._home.clone()
        else:
            m = torch.as_tensor(mask, device=self.sim.device).bool()
            self.sim.reset(m, None)
            self._targets = torch.where(m[:, None], self._home, self._targets)
        return self.get_obs()

    def seed(self, seed):
        np.random.seed(seed)
        random.seed(seed)
        self._random = np.random.RandomState(seed)
        torch.manual_seed(seed)
        if torch.cuda.is_available():
            torch.cuda.manual_seed_all(seed)
        self.action_space.seed(seed)
        self.sim.configure(seed=int(seed) & (2 ** 64 - 1))

    def _physics_step(self, targets, reset_mask=None):
        self.sim.step(targets, reset_mask)

    def step(self, action, reset_mask=None):
        a = self.sim._as_action(action)
        self._targets = a
        self._physics_step(a, reset_mask)
        return None, self.compute_reward(), None, self.get_obs()

    def step_host(self, action, reset_mask=None, host_obs=False):
        """`step` for a host-resident policy (numpy / CPU-tensor actions, the reference's README convention): the kernel
        reads a pinned action batch in place and writes reward / terminated / truncated straight into pinned host arrays.
        Returns (reward, terminated, truncated) as CPU tensors valid on return; observations stay on the device
        (host_obs=True: the kernel also mirrors them into pinned host memory, see `get_obs_host`).  The consumed targets are kept
        ON THE DEVICE by the kernel itself, so `push()`, `control_dofs_position` and masked resets work after a host-fed step."""
        out = self.sim.step_host(action, reset_mask, host_obs=host_obs)
        self._targets = self.sim._last_action           # device tensor written by the step kernel
        return out

    def get_obs_host(self):
        """Pinned host mirrors of the observations written by the last `step_host(..., host_obs=True)`."""
        # alias_outputs: persistent pinned mirrors, overwritten by the next step; default: pinned slots nobody has seen before
        # (the kernel wrote them directly, no copy)
        return {"agent_pos": self.sim.host_obs_agent, "environment_state": self.sim.host_obs_env}

    def compute_reward(self):
        return self.sim.reward

    def get_obs(self):
        return {"agent_pos": self.sim.obs_agent, "environment_state": self.sim.obs_env}

    def get_cams(self):
        raise ValueError("Cameras are not enabled. Set `enable_pixels=True` when creating the environment.")


class FrankaCubePickBatch(CubePickBatch):
    robot_name = "franka"

    @property
    def franka(self):
        return self.robot


class So101CubePickBatch(CubePickBatch):
    robot_name = "so101"

    @property
    def so_101(self):
        return self.robot
