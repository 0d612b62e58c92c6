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


# ---------------------------------------------------------------------------------------------- pileup feature counter, N > 1
def _plp_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from clair3_b200 import pileup_counts as pc, synth_reads as sr
        from oracle import pileup_oracle as po          # stands in for the GPU counter (CPU test of the host logic)
        contig_len, chunk_num = 6000, 5
        rec, ref, rs = sr.random_alignment(contig_len, depth=12, read_len=400, seed=9, origin=0)
        mine = []
        for cid in pc.chunks_for_rank(chunk_num, rank, world):
            a, b = pc.chunk_region(contig_len, cid, chunk_num)
            s, e = pc.counting_region(a, b)
            r = po.clair3_pileup(rec, s, e, ref, rs)
            pos = r["major"][r["cand_cols"]]
            keep = pos[(pos >= a) & (pos < b)]          # candidates of the chunk proper (the widening only feeds the windows)
            mine.append(keep)
        mine = np.concatenate(mine) if mine else np.zeros(0, np.int64)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        if rank == 0:
            whole = po.clair3_pileup(rec, 0, contig_len + 33, ref, rs)
            pos = whole["major"][whole["cand_cols"]]
            q.put((np.concatenate(gathered).tolist(), pos[pos < contig_len].tolist()))
    finally:
        dist.destroy_process_group()


def test_counter_chunks_shard_over_ranks_without_exchange():
    """Contiguous chunk runs per rank, every chunk counted on its widened region: the ranks' candidates, concatenated in rank
    order, are the single-process candidates of the contig (every candidate needs 16 covered columns before it, which the
    33-column widening of preprocess/CreateTensorPileupFromCffi.py:305-311 provides)."""
    from clair3_b200 import pileup_counts as pc
    assert pc.chunks_for_rank(5, 0, 2) == [1, 2, 3] and pc.chunks_for_rank(5, 1, 2) == [4, 5]
    assert sum((pc.chunks_for_rank(7, r, 3) for r in range(3)), []) == list(range(1, 8))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_plp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got, want = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert len(want) > 20 and got == want
