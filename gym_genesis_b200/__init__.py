"""gym_genesis_b200: B200-native batched env.step for the gym-genesis manipulation tasks.

Importing this package registers the same environment ids as the reference
(`gym_genesis/__init__.py:3-37`) with the same default kwargs and `max_episode_steps=200`.
"""
from .gym_api import register

_DEFAULTS = dict(robot="so101", enable_pixels=False, num_envs=10, observation_height=480, observation_width=640,
                 env_spacing=(1.0, 1.0), camera_capture_mode="global", strip_environment_state=True)

for _id, _task in (("gym_genesis/CubePick-v0", "cube_pick"), ("gym_genesis/CubeStack-v0", "cube_stack")):
    register(id=_id, entry_point="gym_genesis_b200.env:GenesisEnv", max_episode_steps=200, nondeterministic=False,
             kwargs=dict(task=_task, **_DEFAULTS))

__version__ = "0.1.0"
