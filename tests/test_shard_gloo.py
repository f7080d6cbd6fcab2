"""world_size-2/3 gloo tests (CPU) of the multi-GPU host logic: rollout sharding, the single score all-gather,
best-of-N selection and winner collection."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dvg_b200 import shard


def test_shard_partition_properties():
    for S in (1, 5, 100, 256, 257):
        for W in (1, 2, 3, 8):
            seen = []
            for r in range(W):
                first, cnt = shard.shard_rollouts(S, W, r)
                seen += list(range(first, first + cnt))
                assert cnt in (S // W, S // W + 1)
            assert seen == list(range(S))
            for g in (0, S - 1):
                r, loc = shard.owner_of(g, S, W)
                first, cnt = shard.shard_rollouts(S, W, r)
                assert first + loc == g and loc < cnt


def test_select_best_handles_nan_and_direction():
    sc = torch.tensor([[1.0, float("nan")], [3.0, 2.0], [2.0, 5.0]])
    assert shard.select_best(sc, True).tolist() == [1, 2]
    assert shard.select_best(sc, False).tolist() == [0, 1]


def _worker(rank, world, port, S, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        scores = torch.rand(S, B, generator=g)                       # identical on every rank
        frames = torch.rand(S, B, 3, 4, generator=g)
        first, cnt = shard.shard_rollouts(S, world, rank)
        allsc = shard.gather_scores(scores[first:first + cnt].clone(), S)
        assert torch.equal(allsc, scores)
        best = shard.select_best(allsc, higher_is_better=True)
        assert torch.equal(best, scores.argmax(0))
        win = shard.gather_winners(frames[first:first + cnt].clone(), best, S)
        want = frames[best, torch.arange(B)]
        assert torch.equal(win, want)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,S", [(2, 10), (2, 7), (3, 8)])
def test_gather_and_select_gloo(world, S):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + world * 10 + S
    procs = [ctx.Process(target=_worker, args=(r, world, port, S, 5, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, "ok") for r in range(world)], res
