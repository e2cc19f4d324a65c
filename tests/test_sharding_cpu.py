"""Host-side logic of the multi-GPU path (batch shard of independent utterances), 2 gloo ranks on CPU.
The per-utterance function here is a stand-in (the real forward needs CUDA); what is tested is the partition,
the absence of cross-rank data dependence and the order-preserving gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vocoder_b200.sharding import gather_batch, shard_bounds, shard_slice, sharded_forward


def test_shard_bounds_partition_every_item_once():
    for n in (0, 1, 5, 32, 64, 255, 256):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                s, e = shard_bounds(n, r, world)
                assert 0 <= s <= e <= n
                seen += list(range(s, e))
            assert seen == list(range(n))
            sizes = [shard_bounds(n, r, world)[1] - shard_bounds(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _fake_generator(mel):  # per-utterance, batch-independent, [B, C, T] -> [B, 1, 4T]
    return mel.mean(dim=1, keepdim=True).repeat_interleave(4, dim=2) * 0.5 + mel[:, :1].repeat_interleave(4, dim=2)


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        mel = torch.randn(n_items, 6, 5)
        full = _fake_generator(mel)
        got = sharded_forward(_fake_generator, mel, gather=True)
        local = sharded_forward(_fake_generator, mel, gather=False)
        s, e = shard_bounds(n_items, rank, world)
        ok = torch.equal(got, full) and (local is None or torch.equal(local, full[s:e]))
        ok = ok and torch.equal(gather_batch(shard_slice(full, rank, world), n_items), full)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [5, 8, 1])
def test_two_rank_gloo_shard_and_gather(n_items):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=10) for _ in range(2))
    assert res == {0: True, 1: True}
