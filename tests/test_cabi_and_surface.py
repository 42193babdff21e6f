"""CPU tier: the C-ABI library loads and exports every symbol include/gg_b200.h declares; host logic of the
gymnasium surface; the product fails loudly without a GPU (no fallback)."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build_cuda()
    from gym_genesis_b200 import cabi
    return cabi.load_library()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "gg_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(gg_[a-z_]+)\s*\(", hdr))
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.gg_version() == 100


def test_model_compiles(franka_model):
    from gym_genesis_b200 import cabi
    g = cabi.compile_gg(franka_model)
    assert (g.n_arm, g.n_fing, g.fing_hinge, g.n_pads) == (7, 2, 0, 10)
    # hand merged into link7: mass 0.735522 + 0.73
    assert abs(g.link_mass[6] - (0.735522 + 0.73)) < 1e-6
    assert list(g.pad_body[:10]) == [1] * 5 + [2] * 5
    assert abs(g.cube_mass - 0.0128) < 1e-9 and abs(g.cube_invweight - 78.125) < 1e-4
    assert g.settle_steps == 1 and g.horizon == 200


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback(franka_model):
    from gym_genesis_b200.sim import BatchedSim, GgError
    with pytest.raises(GgError):
        BatchedSim(franka_model, 4)


def test_registration_matches_reference():
    import gym_genesis  # noqa: F401  (drop-in module path)
    from gym_genesis_b200 import gym_api
    if gym_api.HAVE_GYMNASIUM:
        import gymnasium
        spec = gymnasium.spec("gym_genesis/CubePick-v0")
    else:
        spec = gym_api._REGISTRY["gym_genesis/CubePick-v0"]
        assert "gym_genesis/CubeStack-v0" in gym_api._REGISTRY
    assert spec.max_episode_steps == 200 and spec.nondeterministic is False
    kw = spec.kwargs
    assert kw["task"] == "cube_pick" and kw["robot"] == "so101" and kw["num_envs"] == 10
    assert kw["enable_pixels"] is False and kw["camera_capture_mode"] == "global" and kw["strip_environment_state"] is True
    assert (kw["observation_height"], kw["observation_width"], tuple(kw["env_spacing"])) == (480, 640, (1.0, 1.0))


def test_host_spawn_stream_matches_reference_draw_order():
    """Full reset draws all x then all y from RandomState(seed) (franka/cube_pick.py:90-91)."""
    from gym_genesis_b200.tasks.cube_pick import CubePickBatch
    t = CubePickBatch.__new__(CubePickBatch)
    from gym_genesis_b200.model import franka_cube_pick_model
    t.model, t.num_envs, t._random = franka_cube_pick_model(), 5, np.random.RandomState(42)
    xy = t._spawn_xy()
    r = np.random.RandomState(42)
    x = r.uniform(0.45, 0.80, size=(5,)); y = r.uniform(-0.25, 0.25, size=(5,))
    assert xy.dtype == np.float32 and np.array_equal(xy, np.stack([x, y], 1).astype(np.float32))
