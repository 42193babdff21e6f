"""Drop-in module path: `import gym_genesis` registers gym_genesis/CubePick-v0 on the B200 backend."""
import gym_genesis_b200  # noqa: F401  (registers the env ids)
from gym_genesis_b200 import __version__  # noqa: F401
