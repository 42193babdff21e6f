"""Batched CubeStack task on the B200 simulator.

Mirrors `FrankaCubeStackEmptyBatch` (`gym_genesis/tasks/franka/cube_stack.py:27-240`): Panda + cube_1 (picked) + cube_2
(target) + three distractor cubes on the ground plane; observation `agent_pos` (9) and `environment_state` (14) =
[cube_1 pos, cube_1 quat, eef - cube_1, |.|, cube_2 pos] (`:198-216`); reward = (|xy_1 - xy_2| < 0.05) & (z_1 - z_2 > 0.03)
(`:188-196`); reset draws, per cube, all x then all y from the host RandomState (`:123-145`), sets the home pose, holds it and
runs one settle step (`:147-163`).  The reference maps `CubeStack-v0` + franka to a kitchen-scene variant whose extra
geometry is visual / static (env.py:115); the physics of that task is this scene.
"""
from __future__ import annotations

import numpy as np
import torch

from ..gym_api import spaces
from ..model import get_model
from ..sim import BatchedStackSim
from .cube_pick import CubePickBatch, _Link, _Robot, _Scene

AGENT_DIM = 9       # franka/cube_stack.py:18
ENV_DIM = 14        # franka/cube_stack.py:19


class _StackCube:
    """Handle of object k (`cube_1`, `cube_2`, `distractor_cubes[i]` of the reference): poses read from the SoA state."""

    def __init__(self, task, k):
        self._task, self._k = task, k

    def _rows(self, a, b):
        sim = self._task.sim
        r0 = 2 * sim.nu + 13 * self._k
        return sim.state[r0 + a:r0 + b, : sim.n_envs].t().contiguous()

    def get_pos(self):
        return self._rows(0, 3)

    def get_quat(self):
        return self._rows(3, 7)

    def get_vel(self):
        return self._rows(7, 10)


class _EefLink(_Link):
    """End-effector handle of a task whose agent_pos is NOT the eef pose (SO-101 stack): pose from environment_state's
    eef - cube_1 difference and the cube_1 position."""

    def get_pos(self):
        es = self._task.sim.obs_env
        return es[:, 0:3] + es[:, 7:10]

    def get_quat(self):
        raise NotImplementedError("eef orientation is not part of this task's observations")


class FrankaCubeStackEmptyBatch(CubePickBatch):
    robot_name = "franka"
    scene_name = "empty"

    def __init__(self, enable_pixels, observation_height, observation_width, num_envs, env_spacing, camera_capture_mode,
                 strip_environment_state, device="cuda", auto_reset=False, time_limit=False, solver_iterations=10,
                 solver_tolerance=1e-5, seed=0, first_env_id=0, reset_prob=0.0):
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
        self.model = get_model(self.robot_name, "cube_stack", self.scene_name)
        self.sim = BatchedStackSim(self.model, num_envs, device=device, solver_iterations=solver_iterations,
                                   solver_tolerance=solver_tolerance, auto_reset=auto_reset, time_limit=time_limit,
                                   reset_prob=reset_prob, seed=seed, first_env_id=first_env_id)
        self.agent_dim = self.sim.agent_dim
        self.observation_space = spaces.Dict({
            "agent_pos": spaces.Box(low=-np.inf, high=np.inf, shape=(self.agent_dim,), dtype=np.float32),
            "environment_state": spaces.Box(low=-np.inf, high=np.inf, shape=(ENV_DIM,), dtype=np.float32),
        })
        self.action_space = spaces.Box(low=-1.0, high=1.0, shape=(self.model.nu,), dtype=np.float32)
        self._home = torch.tensor(self.model.home_q, dtype=torch.float32, device=self.sim.device).repeat(num_envs, 1)
        self._targets = self._home.clone()
        self.scene = _Scene(self)
        self.eef = _Link(self) if self.model.agent_mode == "eef" else _EefLink(self)
        self.robot = _Robot(self)
        self.cube_1 = _StackCube(self, 0)
        self.cube_2 = _StackCube(self, 1)
        self.distractor_cubes = [_StackCube(self, k) for k in range(2, self.model.n_obj)]
        self.cube = self.cube_1
        self.motors_dof = np.arange(self.model.n_arm)
        self.fingers_dof = np.arange(self.model.n_arm, self.model.nu)

    @property
    def franka(self):
        return self.robot

    def _spawn_xy(self):
        """Host RandomState stream of the reference (cube_stack.py:123-145): for cube_1, cube_2, then each distractor,
        all x draws followed by all y draws."""
        B = self.num_envs
        out = np.zeros((B, self.model.n_obj, 2), dtype=np.float32)
        for k, (xlo, xhi, ylo, yhi) in enumerate(self.model.obj_spawn):
            out[:, k, 0] = self._random.uniform(xlo, xhi, size=(B,))
            out[:, k, 1] = self._random.uniform(ylo, yhi, size=(B,))
        return out



class FrankaCubeStackKitchenBatch(FrankaCubeStackEmptyBatch):
    """`FrankaCubeStackKitchenBatch` (`gym_genesis/tasks/franka/cube_stack_kitchen_batch.py:26-150`): the variant the reference wires
    to `CubeStack-v0` + franka (`env.py:115`) -- Panda at (-0.5, 0, 0.7) with morph scale 0.6, cubes on the kitchen island
    (`tasks/utils.py:369-427`), spawn ranges `:71-88`; same protocol, observation and reward as the plane variant."""
    scene_name = "kitchen"


class So101CubeStackBatch(FrankaCubeStackEmptyBatch):
    """SO-101 `CubeStackBatch` (`gym_genesis/tasks/so101/cube_stack_batch.py:25-226`): agent_pos = the six joint
    positions (`:176`), environment_state (14) as in the Franka task (the reference declares ENV_DIM = 10 at `:17` but returns
    14 values, `:181`; the space here states what is returned), reward `:150-160`."""
    robot_name = "so101"

    @property
    def so_101(self):
        return self.robot

    def _spawn_xy(self):
        """Host RandomState stream of `cube_stack_batch.py:70-97`: per env, scalar draws x1, y1, x2, y2 repeated until the two
        cubes are >= 0.06 apart; then, per distractor, all x draws followed by all y draws."""
        B = self.num_envs
        out = np.zeros((B, self.model.n_obj, 2), dtype=np.float32)
        (x1lo, x1hi, y1lo, y1hi), (x2lo, x2hi, y2lo, y2hi) = self.model.obj_spawn[0], self.model.obj_spawn[1]
        for b in range(B):
            while True:
                x1 = self._random.uniform(x1lo, x1hi); y1 = self._random.uniform(y1lo, y1hi)
                x2 = self._random.uniform(x2lo, x2hi); y2 = self._random.uniform(y2lo, y2hi)
                if ((x2 - x1) ** 2 + (y2 - y1) ** 2) ** 0.5 >= self.model.stack_min_dist:
                    out[b, 0], out[b, 1] = (x1, y1), (x2, y2)
                    break
        for k in range(2, self.model.n_obj):
            xlo, xhi, ylo, yhi = self.model.obj_spawn[k]
            out[:, k, 0] = self._random.uniform(xlo, xhi, size=(B,))
            out[:, k, 1] = self._random.uniform(ylo, yhi, size=(B,))
        return out


class _UnbatchedStack:
    """Unbatched stacking task (selected by `num_envs=0`, reference env.py:114-120): the batched class at B = 1 with the env
    dimension dropped from observations and reward.  The reference's unbatched resets draw one scalar x / y per cube, which
    consumes the host RandomState exactly like the batched vector draws at B = 1."""
    batch_cls = None

    def __init__(self, enable_pixels, observation_height, observation_width, num_envs, env_spacing, camera_capture_mode,
                 strip_environment_state, **sim_kwargs):
        self._b = self.batch_cls(enable_pixels, observation_height, observation_width, 1, env_spacing, camera_capture_mode,
                                 strip_environment_state, **sim_kwargs)
        self.num_envs = 0
        for name in ("observation_space", "action_space", "scene", "sim", "model", "robot", "eef", "cube", "cube_1", "cube_2",
                     "distractor_cubes", "motors_dof", "fingers_dof", "enable_pixels"):
            setattr(self, name, getattr(self._b, name))

    def seed(self, seed):
        self._b.seed(seed)

    def _squeeze(self, obs):
        return {k: v[0] for k, v in obs.items()}

    def reset(self, mask=None):
        return self._squeeze(self._b.reset(None))

    def step(self, action, reset_mask=None):
        a = torch.as_tensor(action, dtype=torch.float32).reshape(1, -1)
        _, reward, _, obs = self._b.step(a)
        return None, reward[0], None, self._squeeze(obs)

    def compute_reward(self):
        return self._b.compute_reward()[0]

    def get_obs(self):
        return self._squeeze(self._b.get_obs())

    def get_cams(self):
        return self._b.get_cams()


class FrankaCubeStackOne(_UnbatchedStack):
    """`gym_genesis/tasks/franka/cube_stack_one.py`: the unbatched kitchen-island variant (same spawn ranges `:68-85`, home pose and
    gains as `FrankaCubeStackKitchenBatch`)."""
    batch_cls = FrankaCubeStackKitchenBatch

    @property
    def franka(self):
        return self._b.robot


class So101CubeStackOne(_UnbatchedStack):
    """Unbatched SO-101 stacking task (`gym_genesis/tasks/so101/cube_stack.py`, selected by `num_envs=0`, env.py:114,120): the same
    scene, reset stream (scalar draws: one x / y per cube, `:66-93` -- identical to the batched stream at B = 1), reward and
    observations with the env dimension dropped: agent_pos (6,), environment_state (14,), scalar reward."""
    batch_cls = So101CubeStackBatch

    @property
    def so_101(self):
        return self._b.robot
