"""Multi-GPU plumbing for the hot path: contiguous site-range sharding and the one-time weight broadcast.

Every candidate site is independent (no cross-site state, BatchNorm in eval mode), so the path shards with no data-path
collective.  The reference does the same thing with N independent OS processes over disjoint ``.npy``
file lists (``clair3/CallVariantsFromCffiGPU.py:141-156,163-199``); here it is one process per GPU under
``torch.distributed`` and the only collective is a single broadcast of the packed weight image from rank 0 at start-up
(NCCL over NVLink on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def site_range(n_sites: int, rank: int, world: int):
    """Contiguous range [lo, hi) of rank `rank`: preserves per-rank VCF order like the reference's per-GPU file lists.
    The first ``n_sites % world`` ranks take one extra site."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world %d/%d" % (rank, world))
    base, extra = divmod(int(n_sites), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(x, rank: int, world: int):
    lo, hi = site_range(len(x), rank, world)
    return x[lo:hi]


def zeros_like_state_dict(shapes):
    """Placeholder state_dict (right keys/shapes, zero values) for ranks that receive their weights by broadcast."""
    return {k: torch.zeros(tuple(s), dtype=torch.int64 if k.endswith("num_batches_tracked") else torch.float32)
            for k, s in shapes.items()}


class _DevBlob:
    """Minimal __cuda_array_interface__ carrier so torch can alias the library-owned weight image."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3,
                                         "strides": None}


def broadcast_weights(model, src: int = 0, group=None):
    """Broadcast rank `src`'s packed weight image into every rank's model (in place, device to device)."""
    ptr, nbytes = model.weight_blob()
    t = torch.as_tensor(_DevBlob(ptr, nbytes), device=model._device)
    dist.broadcast(t, src=src, group=group)
    torch.cuda.synchronize(model._device)
    return nbytes


def broadcast_state_dict_cpu(state_dict, src: int = 0, group=None):
    """Host-side variant (gloo): broadcast a packed fp32 image of the state_dict; used by the CPU tests and by callers
    that want only rank 0 to read the .pt from disk."""
    keys = sorted(state_dict)
    flat = torch.cat([torch.as_tensor(np.asarray(state_dict[k])).reshape(-1).to(torch.float64) for k in keys])
    dist.broadcast(flat, src=src, group=group)
    out, off = {}, 0
    for k in keys:
        ref = torch.as_tensor(np.asarray(state_dict[k]))
        n = ref.numel()
        out[k] = flat[off:off + n].reshape(ref.shape).to(ref.dtype)
        off += n
    return out


def gather_outputs(y_local, n_sites: int, group=None):
    """All-gather ragged per-rank outputs back into site order (benchmark / parity harness only)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [site_range(n_sites, r, world) for r in range(world)]
    width = y_local.shape[1]
    maxn = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((maxn, width), dtype=y_local.dtype, device=y_local.device)
    pad[:y_local.shape[0]] = y_local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    assert sizes[rank][1] - sizes[rank][0] == y_local.shape[0]
    return torch.cat([b[:hi - lo] for b, (lo, hi) in zip(bufs, sizes)], dim=0)
