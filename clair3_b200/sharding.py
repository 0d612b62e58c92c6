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


def _libnccl():
    """The libnccl torch itself uses (the wheel-bundled one), else the system library."""
    import ctypes
    import importlib.util
    import os
    spec = importlib.util.find_spec("nvidia")
    for root in (list(spec.submodule_search_locations) if spec and spec.submodule_search_locations else []):
        cand = os.path.join(root, "nccl", "lib", "libnccl.so.2")
        if os.path.exists(cand):
            return ctypes.CDLL(cand, mode=ctypes.RTLD_GLOBAL)
    return ctypes.CDLL("libnccl.so.2", mode=ctypes.RTLD_GLOBAL)


class NcclComm:
    """A raw ``ncclComm_t`` over the ranks of the default torch.distributed group, for the C-ABI's ``c3b_bcast_weights``
    (torch does not expose its own communicator).  The 128-byte unique id travels through the existing process group."""

    def __init__(self, device, group=None):
        import ctypes

        class _UniqueId(ctypes.Structure):                 # ncclUniqueId: 128 opaque bytes, passed BY VALUE to ncclCommInitRank
            _fields_ = [("internal", ctypes.c_byte * 128)]

        self._lib = _libnccl()
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        uid = _UniqueId()
        if rank == 0:
            self._lib.ncclGetUniqueId.argtypes = [ctypes.POINTER(_UniqueId)]
            self._lib.ncclGetUniqueId.restype = ctypes.c_int
            rc = self._lib.ncclGetUniqueId(ctypes.byref(uid))
            if rc != 0:
                raise RuntimeError("ncclGetUniqueId failed: %d" % rc)
        t = torch.tensor(list(bytes(uid)), dtype=torch.uint8, device=device if dist.get_backend(group) == "nccl" else "cpu")
        dist.broadcast(t, src=0, group=group)
        ctypes.memmove(ctypes.byref(uid), bytes(t.cpu().tolist()), 128)
        self.comm = ctypes.c_void_p()
        torch.cuda.set_device(device)
        # ncclResult_t ncclCommInitRank(ncclComm_t* comm, int nranks, ncclUniqueId commId /* by value */, int rank)
        self._lib.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, _UniqueId, ctypes.c_int]
        self._lib.ncclCommInitRank.restype = ctypes.c_int
        rc = self._lib.ncclCommInitRank(ctypes.byref(self.comm), world, uid, rank)
        if rc != 0:
            raise RuntimeError("ncclCommInitRank failed: %d" % rc)

    def destroy(self):
        if self.comm:
            self._lib.ncclCommDestroy.argtypes = [__import__("ctypes").c_void_p]
            self._lib.ncclCommDestroy.restype = __import__("ctypes").c_int
            self._lib.ncclCommDestroy(self.comm)
            self.comm = None


def broadcast_weights(model, src: int = 0, group=None, native=True):
    """Broadcast rank `src`'s packed weight images into every rank's model (in place, device to device, one collective per
    image).  ``native``: through the C-ABI's ``c3b_bcast_weights`` on a raw ``ncclComm_t`` (what a C caller does); otherwise
    (or if the raw communicator cannot be created) ``torch.distributed.broadcast`` on tensors aliasing the images.
    Returns (bytes broadcast, "c3b_bcast_weights" | "torch.distributed")."""
    from ._ffi import check, ffi, lib
    total = sum(model.weight_blob(w)[1] for w in (0, 1))
    how = "torch.distributed"
    comm = None
    if native and dist.get_backend(group) == "nccl":
        try:
            comm = NcclComm(model._device, group)
        except Exception:          # no raw libnccl handle: the torch path below is equivalent
            comm = None
    if comm is not None:
        stream = torch.cuda.current_stream(model._device).cuda_stream
        check(lib().c3b_bcast_weights(model._handle, ffi.cast("void *", comm.comm.value), src, ffi.cast("void *", stream)))
        torch.cuda.synchronize(model._device)
        comm.destroy()
        how = "c3b_bcast_weights"
    else:
        for which in (0, 1):
            ptr, nbytes = model.weight_blob(which)
            t = torch.as_tensor(_DevBlob(ptr, nbytes), device=model._device)
            dist.broadcast(t, src=src, group=group)
        torch.cuda.synchronize(model._device)
    if dist.get_rank(group) != src:
        model._by_broadcast = True
    return total, how


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
