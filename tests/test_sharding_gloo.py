"""N>1 host logic on CPU: world_size-2 `gloo` processes run bench.py's sharding path (block partition of
the sequences, no data-path collective, one all-gather of 64-byte per-sequence records, max-over-ranks
timing) and must reproduce the single-process result."""
import os
import socket
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from android_svo_b200 import sharding, capi


def fake_stats(seq_ids):
    """deterministic per-sequence stats (a function of the sequence id only)"""
    st = np.zeros(len(seq_ids), capi.step_stats_dt)
    ids = np.asarray(list(seq_ids))
    st["n_tracked"] = 100 + ids % 7
    st["n_matched"] = 90 + ids % 11
    st["n_seeds_updated"] = 500 + ids % 13
    st["n_seeds_converged"] = ids % 5
    st["align_iters"] = 6 + ids % 3
    perr = np.stack([1e-3 * (1 + ids % 4), 2e-3 * (1 + ids % 3)], 1)
    return st, perr


def worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        usable, per, rng = sharding.shard(total, rank, world)
        st, perr = fake_stats(rng)
        rec = sharding.gather_records(sharding.make_records(list(rng), st, perr), world)
        t = sharding.max_over_ranks(10.0 + rank, world)
        q.put((rank, usable, per, rec, t))
    finally:
        dist.destroy_process_group()


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("total", [64, 65])
def test_block_partition_and_gather_world2(total):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    usable = total - total % world
    st, perr = fake_stats(range(usable))
    expect = sharding.make_records(list(range(usable)), st, perr)
    for rank, u, per, rec, t in res:
        assert u == usable and per == usable // world
        assert rec.shape == (usable, len(sharding.RECORD_FIELDS)) and rec.dtype == np.float64
        assert np.array_equal(rec, expect), "gathered records differ from the single-process result"
        assert t == 11.0                       # max over ranks
    assert sharding.summarize(res[0][3]) == sharding.summarize(expect)
    assert sharding.RECORD_BYTES == 64


def test_shard_covers_everything_once():
    for total in (4096, 4095, 7):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                usable, per, rng = sharding.shard(total, r, world)
                seen += list(rng)
                assert len(rng) == per
            assert seen == list(range(total - total % world))
