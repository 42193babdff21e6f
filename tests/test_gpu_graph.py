"""GPU tier: the step entry points are capturable in a CUDA graph (no allocation, no host synchronisation, every pointer
caller-owned), SURVEY.md section 8b/8d -- replaying the captured graph advances the simulation bit-identically to eager calls."""
import numpy as np
import pytest
import torch

from gym_genesis_b200.model import FRANKA_HOME
from helpers import f32

pytestmark = pytest.mark.gpu


def _targets(n, steps, seed):
    rng = np.random.default_rng(seed)
    t = np.tile(FRANKA_HOME, (steps, n, 1)) + rng.uniform(-0.25, 0.25, (steps, n, 9))
    t[:, :, 7:] = rng.uniform(-0.02, 0.04, (steps, n, 2))
    return torch.as_tensor(f32(t), device="cuda")


def _run(sim, acts, graph):
    buf = torch.zeros_like(acts[0])
    out = []
    if graph:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            buf.copy_(acts[0])
            state0 = sim.state.clone()
            sim.step(buf)                                     # warm-up on the side stream, then restore
            sim.state.copy_(state0)
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            sim.step(buf)
        sim.state.copy_(state0)
        for k in range(acts.shape[0]):
            buf.copy_(acts[k])
            g.replay()
            out.append((sim.obs_env.clone(), sim.reward.clone()))
    else:
        for k in range(acts.shape[0]):
            buf.copy_(acts[k])
            sim.step(buf)
            out.append((sim.obs_env.clone(), sim.reward.clone()))
    torch.cuda.synchronize()
    return out, sim.get_state()


@pytest.mark.parametrize("scene", ["pick", "stack"])
def test_step_is_graph_capturable_and_bit_identical(scene):
    from gym_genesis_b200.model import franka_cube_pick_model, franka_cube_stack_model
    from gym_genesis_b200.sim import BatchedSim, BatchedStackSim
    n, steps = 1024, 12
    acts = _targets(n, steps, 3)
    res = []
    for graph in (False, True):
        if scene == "pick":
            sim = BatchedSim(franka_cube_pick_model(), n, seed=5)
        else:
            sim = BatchedStackSim(franka_cube_stack_model(), n, seed=5)
        sim.reset(None, None)                                 # device Philox spawn: identical in both runs
        res.append(_run(sim, acts, graph))
        sim.close()
    (oa, sa), (ob, sb) = res
    for (ea, ra), (eb, rb) in zip(oa, ob):
        assert torch.equal(ea, eb) and torch.equal(ra, rb)
    for ta, tb in zip(sa, sb):
        assert torch.equal(ta, tb)
