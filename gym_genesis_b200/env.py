"""`GenesisEnv`: the gymnasium facade of the reference (`gym_genesis/env.py:13-125`) on the B200 simulator.

Same constructor keywords, same `reset` / `step` / `push` / `get_*` surface.  Differences are the
ones SURVEY.md Appendix D decides: all five step outputs are torch tensors on the sim device
(`numpy_flags=True` gives host numpy `terminated` / `truncated` exactly as env.py:64-65 intends;
`host_io=True` serves a CPU-resident policy: host actions in, host reward / flags out through pinned zero-copy buffers),
and `get_robot()` / `get_cube()` return the task's handles generically.
"""
from __future__ import annotations

import numpy as np
import torch

from .gym_api import gym
from .tasks import (FrankaCubePickBatch, FrankaCubeStackEmptyBatch, FrankaCubeStackKitchenBatch, FrankaCubeStackOne, So101CubePickBatch,
                    So101CubeStackBatch, So101CubeStackOne)


class GenesisEnv(gym.Env):
    metadata = {"render_modes": ["rgb_array"], "render_fps": 50}

    def __init__(self, task, robot="so101", enable_pixels=False, observation_height=480, observation_width=640, num_envs=1,
                 env_spacing=(1.0, 1.0), render_mode=None, camera_capture_mode="per_env", strip_environment_state=True,
                 numpy_flags=False, host_io=False, host_obs=False, **sim_kwargs):
        super().__init__()
        self.task = task
        self.robot = robot
        self.enable_pixels = enable_pixels
        self.observation_height = observation_height
        self.observation_width = observation_width
        self.num_envs = num_envs
        self.env_spacing = env_spacing
        self.render_mode = render_mode
        self.camera_capture_mode = camera_capture_mode
        self.strip_environment_state = strip_environment_state
        self.numpy_flags = numpy_flags
        self.host_io = host_io
        self.host_obs = host_obs          # host_io only: observations are returned as CPU tensors too (pinned mirrors written by the kernel)
        self._sim_kwargs = sim_kwargs
        self._env = self._make_env_task(self.task)
        self.observation_space = self._env.observation_space
        self.action_space = self._env.action_space
        self.scene = self._env.scene

    def reset(self, seed=None, options=None):
        super().reset(seed=seed)
        if seed is not None:
            self._env.seed(seed)
        mask = None if not options else options.get("reset_mask")
        observation = self._env.reset(mask)
        return observation, {"is_success": [False] * self.num_envs}

    def push(self):
        self._env.scene.step()

    def step(self, action, reset_mask=None):
        if self.host_io:
            # host-resident policy loop (the reference's numpy convention): zero-copy action read, reward / flags land in
            # pinned host memory, one stream synchronisation per step; observations stay on the device
            if self.host_obs:
                reward, terminated, truncated = self._env.step_host(action, reset_mask, host_obs=True)
                observation = self._env.get_obs_host()
            else:
                reward, terminated, truncated = self._env.step_host(action, reset_mask)
                observation = self._env.get_obs()
            if self.numpy_flags:
                terminated, truncated = terminated.numpy(), truncated.numpy()
            return observation, reward, terminated, truncated, {"is_success": terminated}
        _, reward, _, observation = self._env.step(action, reset_mask)
        sim = self._env.sim
        is_success = sim.terminated                 # == (reward == 1), computed in-kernel with the strict z > 0.1f test
        terminated = is_success
        truncated = sim.truncated                   # all False unless time_limit=True was requested
        if self.numpy_flags:
            terminated = terminated.cpu().numpy().astype(bool)
            truncated = truncated.cpu().numpy().astype(bool)
        return observation, reward, terminated, truncated, {"is_success": is_success}

    def save_video(self, save_video=False, file_name="episode.mp4", fps=60):
        if self.enable_pixels and save_video:
            raise NotImplementedError("rendering is out of scope for the state-only path")

    def close(self):
        self._env.sim.close()

    def get_cube(self):
        return self._env.cube

    def get_obs(self):
        return self._env.get_obs()

    def get_robot(self):
        return self._env.robot

    def get_cams(self):
        return self._env.get_cams()

    def render(self):
        return None

    def _make_env_task(self, task_name):
        common_kwargs = dict(enable_pixels=self.enable_pixels, observation_height=self.observation_height,
                             observation_width=self.observation_width, num_envs=self.num_envs, env_spacing=self.env_spacing,
                             camera_capture_mode=self.camera_capture_mode, strip_environment_state=self.strip_environment_state)
        common_kwargs.update(self._sim_kwargs)
        # the reference map (env.py:111-118): ("franka", "cube_stack") is the kitchen-island variant (scaled Panda, cubes on the
        # island top); `scene="empty"` selects FrankaCubeStackEmptyBatch, the plane variant the reference ships but does not wire
        # (SURVEY.md 8f-1).  ("so101", "cube_stack") is CubeStackBatch (8f-3), num_envs = 0 its unbatched form.
        scene = common_kwargs.pop("scene", "kitchen")
        if scene not in ("kitchen", "empty"):
            raise ValueError(f"scene must be 'kitchen' or 'empty', got {scene!r}")
        task_map = {
            ("franka", "cube_pick", True): FrankaCubePickBatch,
            ("so101", "cube_pick", True): So101CubePickBatch,
            ("franka", "cube_stack", True): FrankaCubeStackKitchenBatch if scene == "kitchen" else FrankaCubeStackEmptyBatch,
            ("so101", "cube_stack", True): So101CubeStackBatch,
            ("so101", "cube_stack", False): So101CubeStackOne,       # num_envs = 0: unbatched shapes (env.py:114)
            ("franka", "cube_stack", False): FrankaCubeStackOne,     # env.py:116
        }
        key = (self.robot, task_name, self.num_envs > 0)
        if key not in task_map:
            raise NotImplementedError(key)
        return task_map[key](**common_kwargs)
