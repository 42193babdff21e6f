from .cube_pick import CubePickBatch, FrankaCubePickBatch, So101CubePickBatch  # noqa: F401
