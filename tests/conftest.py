import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def franka_model():
    from gym_genesis_b200.model import franka_cube_pick_model
    return franka_cube_pick_model()


@pytest.fixture(scope="session")
def oracle64(franka_model):
    import numpy as np
    from oracle.binding import Oracle, build
    build()
    return Oracle(franka_model, np.float64)


@pytest.fixture(scope="session")
def oracle32(franka_model):
    import numpy as np
    from oracle.binding import Oracle, build
    build()
    return Oracle(franka_model, np.float32)


@pytest.fixture(scope="session")
def hostsim():
    """Host build of the per-env device header (test harness only)."""
    import ctypes as C
    d = os.path.join(ROOT, "tests", "hostsim")
    subprocess.check_call(["make", "-s", "-C", d])
    return C.CDLL(os.path.join(d, "libgg_hostsim.so"))


@pytest.fixture(scope="session")
def so101_model():
    from gym_genesis_b200.model_so101 import ISLAND_TOP_Z, so101_cube_pick_model
    return so101_cube_pick_model(success_height=ISLAND_TOP_Z + 0.1)


@pytest.fixture(scope="session")
def so101_oracle64(so101_model):
    import numpy as np
    from oracle.binding import Oracle, build
    build()
    return Oracle(so101_model, np.float64)
