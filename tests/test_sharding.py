"""world_size-2 gloo tests (CPU) of the N>1 host logic: site-range sharding, weight broadcast, ordered gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from clair3_b200 import sharding, synth


def test_site_range_partition_properties():
    for n in (0, 1, 7, 8, 1000, 1025):
        for world in (1, 2, 3, 8):
            spans = [sharding.site_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b and c <= d
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.site_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_sites, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sd0 = synth.pileup_state_dict(False, seed=5)
        # rank 0 owns the checkpoint, the others start from zero placeholders and receive it by broadcast
        sd = sd0 if rank == 0 else {k: np.zeros_like(v) for k, v in sd0.items()}
        got = sharding.broadcast_state_dict_cpu(sd, src=0)
        ok_w = all(np.array_equal(got[k].numpy(), sd0[k]) for k in sd0)
        # shard a batch, "process" it (row checksum stands in for the forward), gather back in site order
        x = synth.pileup_inputs(n_sites, seed=5)
        mine = sharding.shard_batch(x, rank, world)
        y_local = torch.from_numpy(mine.reshape(len(mine), -1).sum(1, keepdims=True).astype(np.float32))
        y = sharding.gather_outputs(y_local, n_sites)
        ref = x.reshape(n_sites, -1).sum(1, keepdims=True).astype(np.float32)
        q.put((rank, ok_w, bool(np.array_equal(y.numpy(), ref)), len(mine)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_sites", [101, 64])
def test_two_rank_broadcast_shard_gather(n_sites):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_sites, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] and r[2] for r in res)
    assert sum(r[3] for r in res) == n_sites
