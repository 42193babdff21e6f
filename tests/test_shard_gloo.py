"""CPU tier, world_size 2 on gloo: the host-side sharding logic and the statistics all-gather (the only collective)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gym_genesis_b200.shard import env_shard, gather_stats, summarize


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = env_shard(total_envs=131072)
    local = torch.arange(8, dtype=torch.float64) * (rank + 1)
    allst = gather_stats(local)
    q.put((rank, sh.first_env_id, sh.n_local, allst.tolist()))
    dist.destroy_process_group()


def test_two_rank_sharding_and_stats_gather():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[0], r[1], r[2]) for r in res] == [(0, 0, 65536), (1, 65536, 65536)]      # contiguous, disjoint, covering
    expect = [[float(i) for i in range(8)], [2.0 * i for i in range(8)]]
    assert res[0][3] == expect and res[1][3] == expect                                  # every rank sees every rank's vector


def test_shard_validation_and_summary():
    assert env_shard(envs_per_rank=16384, rank=3, world=8).first_env_id == 49152
    with pytest.raises(ValueError):
        env_shard(total_envs=10, rank=0, world=4)
    with pytest.raises(ValueError):
        env_shard(total_envs=8, envs_per_rank=2, rank=0, world=4)
    s = summarize(torch.tensor([[100, 4, 2, 2, 400, 0, 50, 400], [100, 4, 4, 4, 200, 0, 50, 400]], dtype=torch.float64))
    assert s["episodes"] == 8 and s["success_rate"] == 0.75 and s["mean_episode_length"] == 75.0
    assert gather_stats(torch.ones(8, dtype=torch.float64)).shape == (1, 8)
