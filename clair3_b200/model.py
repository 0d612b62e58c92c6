"""Drop-in ``Clair3_P`` / ``Clair3_F`` for the reference's callers, backed by the sm_100a kernels.

Mirrors the module protocol the reference uses (HKU-BAL/Clair3 paths):

    m = Clair3_P|Clair3_F(add_indel_length, predict=True, input_channels)   clair3/CallVariantsFromCffi.py:230-243
    m.to(device); m.eval(); m.load_state_dict(state_dict)                    clair3/CallVariantsFromCffi.py:19-28,246-248
    Y = m(X)    # X int8/int32/float tensor [B,33,18] or [B,D,33,C]; Y float32 [B,24|90] on X's device   :48-52

Constructor arguments, state_dict keys (strict), output head order and dtype are the reference's
(``clair3/model.py:58-161`` and ``:282-416``).  Everything numeric happens in ``libclair3b200.so``; there is no
PyTorch or CPU implementation behind these classes, and constructing one without a B200 raises.
"""
from __future__ import annotations

import numpy as np
import torch

from ._ffi import CONSTANTS as K
from ._ffi import C3BError, check, ffi, lib

_DT = {torch.int8: K["C3B_DT_I8"], torch.int32: K["C3B_DT_I32"], torch.float32: K["C3B_DT_F32"]}


class _C3BModule:
    _kind = None
    _default_channels = None

    def __init__(self, add_indel_length=False, predict=False, input_channels=None):
        self.add_indel_length = bool(add_indel_length)
        self.predict = bool(predict)
        self.input_channels = int(input_channels) if input_channels is not None else self._default_channels
        self.output_label_split = [21, 3, 33, 33]
        self.training = False
        self._device = None
        self._handle = None
        self._state = None
        self._options = {}

    # ---- torch.nn.Module-shaped surface used by the reference callers
    def to(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise C3BError("clair3_b200 runs only on a CUDA (sm_100a) device; there is no CPU path "
                           "(requested device: %s)" % device)
        index = device.index if device.index is not None else torch.cuda.current_device()
        if self._handle is not None and self._device is not None and self._device.index == index:
            return self
        self._release()
        out = ffi.new("c3b_model **")
        check(lib().c3b_create(out, self._kind, self.input_channels, int(self.add_indel_length), index))
        self._handle = out[0]
        self._device = torch.device("cuda", index)
        for k, v in self._options.items():
            check(lib().c3b_set_option(self._handle, k.encode(), int(v)))
        if self._state is not None:
            self._upload(self._state)
        return self

    def cuda(self, device=None):
        return self.to(torch.device("cuda", device if device is not None else torch.cuda.current_device()))

    def eval(self):
        self.training = False
        return self

    def train(self, mode=True):
        if mode:
            raise C3BError("clair3_b200 implements the inference forward only (reference training is clair3/Train.py)")
        return self

    def state_dict(self):
        return dict(self._state or {})

    def load_state_dict(self, state_dict, strict=True):
        sd = {}
        for k, v in state_dict.items():
            t = v.detach().cpu() if isinstance(v, torch.Tensor) else torch.as_tensor(np.asarray(v))
            sd[k] = t
        self._state = sd
        if self._handle is None:
            if torch.cuda.is_available():
                self.to(torch.device("cuda"))
            else:
                raise C3BError("no CUDA device: clair3_b200 has no CPU fallback")
        else:
            self._upload(sd)
        return self

    def set_option(self, name, value):
        """Kernel options: precision (0 bf16 tensor cores | 1 fp32 debug), chunk_sites, lstm_tile."""
        self._options[name] = int(value)
        if self._handle is not None:
            check(lib().c3b_set_option(self._handle, name.encode(), int(value)))
        return self

    def _upload(self, sd):
        L = lib()
        for key, t in sd.items():
            if t.dtype == torch.int64:
                arr = t.contiguous().numpy()
                dt = K["C3B_DT_I64"]
            else:
                arr = t.to(torch.float32).contiguous().numpy()
                dt = K["C3B_DT_F32"]
            dims = list(t.shape)                         # 0-d tensors (num_batches_tracked) keep ndim = 0
            shape = ffi.new("int64_t[]", dims or [0])
            check(L.c3b_set_param(self._handle, key.encode(), ffi.cast("void *", arr.ctypes.data), dt, shape, len(dims)))
        check(L.c3b_finalize(self._handle))

    @property
    def out_dim(self):
        return 90 if self.add_indel_length else 24

    def forward(self, x):
        if self._handle is None:
            raise C3BError("model has no device/weights yet: call .to(device) and .load_state_dict() first")
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(x)
        want_nd = 3 if self._kind == K["C3B_PILEUP"] else 4
        if x.ndim != want_nd or x.shape[-1] != self.input_channels or x.shape[-2] != 33:
            raise C3BError("expected input [B,%s33,%d], got %s" % ("" if want_nd == 3 else "depth,", self.input_channels,
                                                                  tuple(x.shape)))
        if x.dtype not in _DT:
            x = x.to(torch.int32) if not x.dtype.is_floating_point else x.to(torch.float32)
        x = x.contiguous()
        batch = x.shape[0]
        depth = x.shape[1] if want_nd == 4 else 0
        on_dev = x.device.type == "cuda"
        if on_dev and x.device.index != self._device.index:
            raise C3BError("input is on %s but the model is on %s" % (x.device, self._device))
        y = torch.empty((batch, self.out_dim), dtype=torch.float32, device=x.device if on_dev else "cpu")
        if batch == 0:
            return self._split(y)
        stream = torch.cuda.current_stream(self._device).cuda_stream
        check(lib().c3b_forward(self._handle, ffi.cast("void *", x.data_ptr()), _DT[x.dtype], int(on_dev), batch, depth,
                                ffi.cast("float *", y.data_ptr()), int(on_dev), ffi.cast("void *", stream)))
        return self._split(y)

    __call__ = forward

    def forward_async(self, x_host, y_host):
        """Stream-ordered forward on PINNED host tensors (H2D -> kernels -> D2H on the current CUDA stream, no host
        synchronisation): the double-buffered caller of SURVEY.md §8f N1.  The caller synchronises the stream before
        reading ``y_host`` and must keep both tensors alive until then."""
        if not (x_host.is_pinned() and y_host.is_pinned()):
            raise C3BError("forward_async needs pinned host tensors")
        if x_host.dtype not in _DT or not x_host.is_contiguous() or y_host.dtype != torch.float32:
            raise C3BError("forward_async: x must be contiguous int8/int32/float32, y float32")
        batch = x_host.shape[0]
        depth = x_host.shape[1] if x_host.ndim == 4 else 0
        if tuple(y_host.shape) != (batch, self.out_dim):
            raise C3BError("forward_async: y must be [batch, %d]" % self.out_dim)
        self.set_option("host_async", 1)
        try:
            stream = torch.cuda.current_stream(self._device).cuda_stream
            check(lib().c3b_forward(self._handle, ffi.cast("void *", x_host.data_ptr()), _DT[x_host.dtype], 0, batch, depth,
                                    ffi.cast("float *", y_host.data_ptr()), 0, ffi.cast("void *", stream)))
        finally:
            self.set_option("host_async", 0)
        return y_host

    def _split(self, y):
        if self.predict:
            return y
        sizes = self.output_label_split[:4 if self.add_indel_length else 2]
        return list(torch.split(y, sizes, dim=1))

    # ---- extras
    def tap(self, name):
        """Intermediate activation of the last forward as float32 numpy (debug/parity)."""
        cnt = ffi.new("int64_t *", 0)
        lib().c3b_get_tap(self._handle, name.encode(), ffi.NULL, cnt)
        n = int(cnt[0])
        if n <= 0:
            raise C3BError(ffi.string(lib().c3b_last_error()).decode())
        out = np.empty(n, dtype=np.float32)
        cnt[0] = n
        check(lib().c3b_get_tap(self._handle, name.encode(), ffi.cast("float *", out.ctypes.data), cnt))
        return out[:int(cnt[0])]

    def weight_blob(self):
        """(device_ptr, nbytes) of the packed weight image (the broadcast unit of clair3_b200.sharding)."""
        p = ffi.new("void **")
        n = ffi.new("size_t *")
        check(lib().c3b_weight_blob(self._handle, p, n))
        return int(ffi.cast("uintptr_t", p[0])), int(n[0])

    @property
    def launch_count(self):
        return int(lib().c3b_launch_count(self._handle)) if self._handle is not None else 0

    def _release(self):
        if self._handle is not None:
            lib().c3b_destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass


class Clair3_P(_C3BModule):
    """Pileup network forward (reference: clair3/model.py:58-161)."""
    _kind = K["C3B_PILEUP"]
    _default_channels = 18      # shared/param_p.py:31-36


class Clair3_F(_C3BModule):
    """Full-alignment network forward (reference: clair3/model.py:282-416)."""
    _kind = K["C3B_FULL_ALIGNMENT"]
    _default_channels = 8       # shared/param_f.py:23-30
