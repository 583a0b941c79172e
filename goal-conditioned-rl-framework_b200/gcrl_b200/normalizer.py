"""RunningNormalizer with device-resident state (reference src/utils.py:68-117)."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import yaml

from . import _lib
from ._lib import check, lib, np_ptr, vp


class RunningNormalizer:
    """Same constructor, methods and YAML format as the reference class; the running
    (mean, var, count) live in float64 on the GPU and ``update`` / ``normalize`` run the
    CUDA kernels in csrc/normalizer.cu.  ``mean`` / ``var`` / ``count`` read back."""

    def __init__(self, size, clip_range=5.0, eps=1e-8, device=0):
        _lib.require_cuda()
        self.size = int(size)
        self.device_index = int(device)
        h = vp()
        check(lib.gcrl_norm_create(C.byref(h), self.device_index, self.size, float(clip_range),
                                   float(eps)))
        self._h = h
        self._clip = float(clip_range)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib.gcrl_norm_destroy(h)
            self._h = None

    # -- state ------------------------------------------------------------------------
    def _state(self):
        mean = np.empty(self.size, np.float64)
        var = np.empty(self.size, np.float64)
        cnt, clip = C.c_double(), C.c_double()
        check(lib.gcrl_norm_get_state(self._h, np_ptr(mean), np_ptr(var), C.byref(cnt),
                                      C.byref(clip), self._stream()))
        return mean, var, cnt.value, clip.value

    def _stream(self):
        return _lib.current_stream(self.device_index)

    mean = property(lambda self: self._state()[0])
    var = property(lambda self: self._state()[1])
    count = property(lambda self: self._state()[2])

    @property
    def clip_range(self):
        return self._clip

    def set_state(self, mean, var, count, clip_range=None):
        mean = np.ascontiguousarray(mean, np.float64).reshape(self.size)
        var = np.ascontiguousarray(var, np.float64).reshape(self.size)
        if clip_range is not None:
            self._clip = float(clip_range)
        check(lib.gcrl_norm_set_state(self._h, np_ptr(mean), np_ptr(var), float(count),
                                      self._clip, self._stream()))

    # -- reference API ------------------------------------------------------------------
    @staticmethod
    def _as_rows(x, size):
        x = np.asarray(x)
        if x.dtype != np.float64 and x.dtype != np.float32:
            x = x.astype(np.float64)
        shape = x.shape
        x = np.ascontiguousarray(x).reshape(-1, size)
        return x, shape

    def update(self, x):                                   # src/utils.py:75-80
        x, _ = self._as_rows(x, self.size)
        if self._gather is not None:
            return self.update_moments(self._gather(self.batch_moments(x)))
        check(lib.gcrl_norm_update(self._h, np_ptr(x), x.shape[0], int(x.dtype == np.float64),
                                   self._stream()))

    # -- data parallel: every rank folds all ranks' batch moments, in rank order --------
    _gather = None

    def enable_data_parallel(self, process_group=None, gather=None):
        """From now on ``update(x)`` is collective: the ranks' batch moments (2 * dim + 1 float64 numbers each)
        are all-gathered and merged in rank order, so every replica of the running statistics stays
        bit-identical and equals a single-process update on the concatenated batch to float64 rounding.
        ``gather(moments [dim, 3]) -> [world, dim, 3]`` overrides the collective (tests)."""
        if gather is None:
            import torch
            import torch.distributed as dist

            def gather(m):
                nccl = dist.get_backend(process_group) == "nccl"
                t = torch.from_numpy(m)
                if nccl:
                    t = t.cuda(self.device_index)
                out = [torch.empty_like(t) for _ in range(dist.get_world_size(process_group))]
                dist.all_gather(out, t, group=process_group)
                return np.stack([o.cpu().numpy() for o in out])
        self._gather = gather

    def batch_moments(self, x):
        """(n, mean, M2) per column of a batch, float64 [dim, 3]; the running state is not touched."""
        x, _ = self._as_rows(x, self.size)
        out = np.empty((self.size, 3), np.float64)
        check(lib.gcrl_norm_batch_moments(self._h, np_ptr(x), x.shape[0], int(x.dtype == np.float64), np_ptr(out),
                                          self._stream()))
        return out

    def update_moments(self, moments):
        """Fold batch moments [parts, dim, 3] (in the given order) into the running state."""
        m = np.ascontiguousarray(moments, np.float64).reshape(-1, self.size, 3)
        check(lib.gcrl_norm_update_moments(self._h, np_ptr(m), m.shape[0], self._stream()))

    def normalize(self, x):                                # src/utils.py:96-98
        x, shape = self._as_rows(x, self.size)
        out = np.empty(x.shape, np.float64)
        check(lib.gcrl_norm_apply(self._h, np_ptr(x), x.shape[0], int(x.dtype == np.float64),
                                  np_ptr(out), self._stream()))
        return out.reshape(shape)

    def save(self, path: str):                             # src/utils.py:100-109
        os.makedirs(os.path.dirname(path), exist_ok=True)
        mean, var, count, clip = self._state()
        with open(path, "w") as f:
            yaml.dump({"mean": mean.tolist(), "var": var.tolist(), "count": float(count),
                       "clip_range": float(clip)}, f)

    def load(self, path: str):                             # src/utils.py:111-117
        with open(path, "r") as f:
            data = yaml.safe_load(f)
        # the reference narrows the loaded statistics to float32 (:114-115)
        mean = np.array(data["mean"], dtype=np.float32)
        var = np.array(data["var"], dtype=np.float32)
        self.set_state(mean, var, float(data["count"]), float(data["clip_range"]))
