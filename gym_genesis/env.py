"""Drop-in for `gym_genesis.env` (reference gym_genesis/env.py): re-exports the B200 facade."""
from gym_genesis_b200.env import GenesisEnv  # noqa: F401
