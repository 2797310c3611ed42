"""N > 1 host logic on CPU: env sharding + the episode-statistics all-reduce over a world_size-2 gloo group, and the
shard-invariance of the Philox reset streams (checked with the host build of the reset code)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from jitterbug_b200.distributed import gather_episode_stats, shard_envs


def test_shard_envs_partition():
    for total in (1, 7, 8, 4096, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard_envs(total, r, world) for r in range(world)]
            assert spans[0][0] == 0
            for (o0, c0), (o1, _) in zip(spans, spans[1:]):
                assert o0 + c0 == o1
            assert spans[-1][0] + spans[-1][1] == total
            counts = [c for _, c in spans]
            assert max(counts) - min(counts) <= 1


def _worker(rank, world, port, total, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        all_returns = rng.uniform(0, 1000, size=total)
        off, cnt = shard_envs(total, rank, world)
        mine = torch.from_numpy(all_returns[off:off + cnt])
        stats = gather_episode_stats(mine)
        # the running-monitor form (rollout.VecJitterbugEnv accumulates these moments on the device)
        moments = (torch.tensor(float(cnt), dtype=torch.float64), mine.sum(), (mine * mine).sum(), mine.min(), mine.max())
        stats_m = gather_episode_stats(None, moments=moments)
        assert stats_m == stats and float(moments[3]) == float(mine.min())      # inputs are not modified
        if rank == 0:
            out.put((stats, float(all_returns.mean()), float(all_returns.std()), float(all_returns.min()),
                     float(all_returns.max())))
    finally:
        dist.destroy_process_group()


def test_gather_episode_stats_gloo_world2():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    total = 1001
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    stats, mean, std, lo, hi = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert stats["count"] == total
    assert stats["mean"] == pytest.approx(mean, rel=1e-12)
    assert stats["std"] == pytest.approx(std, rel=1e-9)
    assert stats["min"] == lo and stats["max"] == hi


def test_reset_streams_are_keyed_by_global_env_id():
    """Two shards (offsets 0 and 5) reproduce the single-shard reset states: same (seed, global id, episode) -> same draw."""
    import hostsim
    full = hostsim.reset_states(seed=123, env_ids=np.arange(10), episode=3, task=4, random_pose=1)
    a = hostsim.reset_states(seed=123, env_ids=np.arange(0, 5), episode=3, task=4, random_pose=1)
    b = hostsim.reset_states(seed=123, env_ids=np.arange(5, 10), episode=3, task=4, random_pose=1)
    np.testing.assert_array_equal(full, np.concatenate([a, b]))
    # distribution sanity (jitterbug.py:609-664): radius in [.05, .2], unit quaternion, tilt axis within ~2.5 % of z
    many = hostsim.reset_states(seed=7, env_ids=np.arange(4000), episode=0, task=4, random_pose=1)
    quat, tgt = many[:, 0:4], many[:, 4:8]
    np.testing.assert_allclose(np.linalg.norm(quat, axis=1), 1, atol=1e-6)
    rad = np.hypot(tgt[:, 0], tgt[:, 1])
    assert rad.min() >= 0.05 - 1e-6 and rad.max() <= 0.2 + 1e-6
    assert abs(rad.mean() - 0.125) < 0.005
    np.testing.assert_allclose(tgt[:, 2] ** 2 + tgt[:, 3] ** 2, 1, atol=1e-6)
    axis_xy = quat[:, 1:3] / np.maximum(np.abs(quat[:, 3:4]), 1e-9)
    assert np.abs(axis_xy).max() <= 0.0251
    other = hostsim.reset_states(seed=8, env_ids=np.arange(10), episode=3, task=4, random_pose=1)
    assert not np.array_equal(other, full)
