"""Device-side trajectory store and dataset export for the data collectors.

The reference collectors (`examples/franka/pick_cube_state.py:86-113`) append `obs[...].detach().cpu().numpy()`, the action
and the reward to host lists after EVERY env.step (a blocking device->host copy per step and per array), stack them into
(T, B, dim) arrays ("FIXME: system ram crash if B is too big", `:96`), then walk all envs and hand the frames of every env
whose reward was ever > 0 to `LeRobotDataset.add_frame` / `save_episode`.

Here the frames stay on the GPU: `TrajectoryRecorder.record(t, action)` is one copy kernel (`gg_record`) per step with no
host synchronisation; `export_successful()` compacts the successful envs on the device (`gg_export_successful`) and brings
them to the host with ONE copy, episode-major, in the order the reference would have saved them.
`write_lerobot_dataset` writes those episodes in the LeRobotDataset v2.1 on-disk layout (parquet + meta jsonl) with the
reference's feature names; `lerobot` itself is not a dependency.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from typing import Optional

import numpy as np
import torch

from .sim import BatchedSim, _ptr


class TrajectoryRecorder:
    """(horizon, n_envs, F) float32 frame store on the simulator's device; frame = [agent_pos | environment_state | action | reward]."""

    def __init__(self, sim: BatchedSim, horizon: int):
        if horizon <= 0:
            raise ValueError("horizon must be positive")
        self.sim, self.horizon = sim, int(horizon)
        # the stacking scenes (BatchedStackSim) have their own entry points and a 14-value environment_state
        stack = hasattr(sim, "n_obj")
        self._f_dim, self._f_rec, self._f_exp = ((sim.lib.gg_stack_frame_dim, sim.lib.gg_stack_record, sim.lib.gg_stack_export_successful)
                                                 if stack else (sim.lib.gg_frame_dim, sim.lib.gg_record, sim.lib.gg_export_successful))
        self.frame_dim = int(self._f_dim(sim._h))
        dev = sim.device
        self.traj = torch.zeros((self.horizon, sim.n_envs, self.frame_dim), dtype=torch.float32, device=dev)
        self.ever = torch.zeros(sim.n_envs, dtype=torch.int32, device=dev)
        self._out = None
        self._host = None
        a, e, u = sim.agent_dim, (sim.env_dim if stack else 11), sim.nu
        self.slices = {"agent_pos": slice(0, a), "environment_state": slice(a, a + e), "action": slice(a + e, a + e + u),
                       "reward": slice(a + e + u, a + e + u + 1)}

    def begin_episode(self):
        self.ever.zero_()

    def record(self, t: int, action: torch.Tensor):
        """Store frame t: the simulator's current observation / reward buffers (the outputs of the step that consumed `action`)."""
        a = self.sim._as_action(action)
        s = self.sim
        s._check(self._f_rec(s._h, _ptr(self.traj), _ptr(self.ever), int(t), self.horizon, _ptr(a), _ptr(s.obs_agent), _ptr(s.obs_env),
                                 _ptr(s.reward), s._stream()))

    def export_successful(self):
        """-> (frames (n_selected, horizon, F) pinned CPU tensor, env_ids (n_selected,) CPU int32): the envs whose reward was > 0 at
        any recorded step, in env order (the reference's `for b in range(num_envs): if np.any(rewards_arr[:, b] > 0)`).
        `frames` is a view of a pinned buffer that the next export_successful() overwrites."""
        s = self.sim
        if self._out is None:
            self._out = (torch.empty((s.n_envs, self.horizon, self.frame_dim), dtype=torch.float32, device=s.device),
                         torch.empty(s.n_envs, dtype=torch.int32, device=s.device), torch.zeros(1, dtype=torch.int32, device=s.device))
        out, ids, cnt = self._out
        s._check(self._f_exp(s._h, _ptr(self.traj), _ptr(self.ever), self.horizon, _ptr(out), _ptr(ids), _ptr(cnt), s._stream()))
        k = int(cnt.item())                       # the one host sync of the episode
        if self._host is None or self._host.shape[0] < k:
            # pinned landing buffer, grown geometrically and re-used across episodes (page-locking 100s of MB costs ~100 ms)
            self._host = torch.empty((min(s.n_envs, max(2 * k, 64)), self.horizon, self.frame_dim), dtype=torch.float32).pin_memory()
        frames = self._host[:k]
        if k:
            frames.copy_(out[:k], non_blocking=True)
        env_ids = ids[:k].cpu()
        torch.cuda.current_stream(s.device).synchronize()
        return frames, env_ids


# ------------------------------------------------------------------------------------------------
# LeRobotDataset v2.1 on-disk layout ([UPSTREAM-RECALL]: lerobot is not installed here):
#   meta/info.json, meta/tasks.jsonl, meta/episodes.jsonl, data/chunk-000/episode_000000.parquet ...
# One parquet file per episode; columns = the features plus timestamp / frame_index / episode_index / index / task_index.
# ------------------------------------------------------------------------------------------------
class LeRobotWriter:
    CHUNK = 1000

    def __init__(self, root: str, fps: int = 60, robot_type: str = "franka", features: Optional[dict] = None):
        self.root, self.fps, self.robot_type = root, int(fps), robot_type
        self.features = features or {}
        self.n_episodes = 0
        self.n_frames = 0
        self.tasks: dict[str, int] = {}
        self.episodes = []
        os.makedirs(os.path.join(root, "meta"), exist_ok=True)

    def add_episode(self, columns: dict, task: str):
        """columns: {feature name: (T, dim) float32 array}; all share T."""
        import pyarrow as pa
        import pyarrow.parquet as pq
        T = next(iter(columns.values())).shape[0]
        ti = self.tasks.setdefault(task, len(self.tasks))
        ep = self.n_episodes
        cols = {}
        for name, arr in columns.items():
            arr = np.ascontiguousarray(arr, dtype=np.float32)
            if arr.shape[0] != T:
                raise ValueError("all features of an episode need the same number of frames")
            cols[name] = pa.FixedSizeListArray.from_arrays(pa.array(arr.reshape(-1), type=pa.float32()), arr.shape[1])
            self.features.setdefault(name, {"dtype": "float32", "shape": [int(arr.shape[1])], "names": None})
        cols["timestamp"] = pa.array(np.arange(T, dtype=np.float32) / self.fps)
        cols["frame_index"] = pa.array(np.arange(T, dtype=np.int64))
        cols["episode_index"] = pa.array(np.full(T, ep, dtype=np.int64))
        cols["index"] = pa.array(np.arange(self.n_frames, self.n_frames + T, dtype=np.int64))
        cols["task_index"] = pa.array(np.full(T, ti, dtype=np.int64))
        d = os.path.join(self.root, "data", f"chunk-{ep // self.CHUNK:03d}")
        os.makedirs(d, exist_ok=True)
        pq.write_table(pa.table(cols), os.path.join(d, f"episode_{ep:06d}.parquet"))
        self.episodes.append({"episode_index": ep, "tasks": [task], "length": int(T)})
        self.n_episodes += 1
        self.n_frames += T

    def finalize(self):
        feats = dict(self.features)
        for k, dt in (("timestamp", "float32"), ("frame_index", "int64"), ("episode_index", "int64"), ("index", "int64"), ("task_index", "int64")):
            feats[k] = {"dtype": dt, "shape": [1], "names": None}
        info = {"codebase_version": "v2.1", "robot_type": self.robot_type, "total_episodes": self.n_episodes, "total_frames": self.n_frames,
                "total_tasks": len(self.tasks), "total_videos": 0, "total_chunks": (self.n_episodes + self.CHUNK - 1) // self.CHUNK,
                "chunks_size": self.CHUNK, "fps": self.fps, "splits": {"train": f"0:{self.n_episodes}"},
                "data_path": "data/chunk-{episode_chunk:03d}/episode_{episode_index:06d}.parquet", "video_path": None, "features": feats}
        with open(os.path.join(self.root, "meta", "info.json"), "w") as f:
            json.dump(info, f, indent=2)
        with open(os.path.join(self.root, "meta", "tasks.jsonl"), "w") as f:
            for t, i in sorted(self.tasks.items(), key=lambda kv: kv[1]):
                f.write(json.dumps({"task_index": i, "task": t}) + "\n")
        with open(os.path.join(self.root, "meta", "episodes.jsonl"), "w") as f:
            for e in self.episodes:
                f.write(json.dumps(e) + "\n")
        return info


def write_lerobot_dataset(writer: LeRobotWriter, recorder: TrajectoryRecorder, frames, task: str):
    """Hand the exported episodes to the writer with the reference's feature names (pick_cube_state.py:103-108; the collector's
    dataset declares `observation.agent_pos` but adds frames as `observation.state` -- the name used when frames are added is kept)."""
    fr = frames.numpy() if isinstance(frames, torch.Tensor) else np.asarray(frames)
    sl = recorder.slices
    for k in range(fr.shape[0]):
        writer.add_episode({"observation.state": fr[k][:, sl["agent_pos"]], "observation.environment_state": fr[k][:, sl["environment_state"]],
                            "action": fr[k][:, sl["action"]]}, task)
