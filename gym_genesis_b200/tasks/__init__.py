from .cube_pick import CubePickBatch, FrankaCubePickBatch, So101CubePickBatch  # noqa: F401
from .cube_stack import (FrankaCubeStackEmptyBatch, FrankaCubeStackKitchenBatch, FrankaCubeStackOne,  # noqa: F401
                         So101CubeStackBatch, So101CubeStackOne)
