"""Host logic of the fresh-output slabs (gym_genesis_b200/sim.py: _SlabRing): per-step tensor sets are views of K-step slabs; a slab is
split again only after the caller dropped every view of it, otherwise a new slab is allocated -- fresh semantics, no steady-state
allocation.  CPU tensors, no GPU."""
import torch

from gym_genesis_b200.sim import _SlabRing


def _ring(K, counter):
    def make():
        counter[0] += 1
        return (torch.zeros((K, 4)), torch.zeros((2, K, 3), dtype=torch.bool))

    def split(bases):
        A, T, U = bases[0].unbind(0), bases[1][0].unbind(0), bases[1][1].unbind(0)
        return [(A[k], T[k], U[k]) for k in range(len(A))]
    return _SlabRing(make, split)


def test_steady_loop_recycles_two_slabs():
    made = [0]
    ring = _ring(4, made)
    last = None
    for step in range(100):
        a, t, u = ring.next()
        a.fill_(step)
        if last is not None:
            assert last[0][0].item() == step - 1          # the previous step's result is intact while the caller still holds it
        last = (a, t, u)
    assert made[0] == 2 and len(ring.slabs) == 2          # one slab backs the latest results while the other is being handed out


def test_kept_results_force_new_slabs_and_keep_their_values():
    made = [0]
    ring = _ring(4, made)
    kept = []
    for step in range(40):
        a, t, u = ring.next()
        a.fill_(step); t.fill_(step % 2 == 0)
        kept.append((a, t))
    assert made[0] == 10                                  # nothing could be recycled: 40 steps / 4 per slab
    assert len(ring.slabs) <= _SlabRing.MAX_SLABS         # the ring itself stays bounded; older slabs live on through the kept views
    for step, (a, t) in enumerate(kept):
        assert (a == step).all() and bool(t.all()) == (step % 2 == 0)
    assert len({a.data_ptr() for a, _ in kept}) == 40
    del kept, a, t, u
    before = made[0]
    for step in range(40):
        ring.next()
    assert made[0] == before                              # everything was dropped: the ring recycles again


def test_partial_hold_blocks_only_that_slab():
    made = [0]
    ring = _ring(4, made)
    held = ring.next()[0]                                 # one view of the first slab stays alive
    held.fill_(7.0)
    for step in range(64):
        ring.next()[0].fill_(-1.0)
    assert (held == 7.0).all()
    assert made[0] <= 4
