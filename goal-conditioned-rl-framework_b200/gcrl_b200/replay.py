"""ReplayBuffer / PERBuffer over the device-resident transition ring (reference src/buffer.py:8-89)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, lib, np_ptr, vp


def _to_np(x, dtype=np.float32):
    if hasattr(x, "detach"):                 # torch tensor (env.py hands device rows)
        x = x.detach().cpu().numpy()
    return np.asarray(x, dtype=dtype)


class _Ring:
    """Shared storage: ``push`` stages packed rows on the host and commits them to the GPU in one copy
    before the next ``sample`` (or every ``_FLUSH`` rows); deque(maxlen) semantics."""

    _PRIORITIZED = False
    _FLUSH = 4096

    def __init__(self, max_len: int, alpha: float = 0.0, *, device=0):
        _lib.require_cuda()
        self.max_len = int(max_len)
        self.alpha = alpha
        self.device_index = int(device)
        self.device = f"cuda:{self.device_index}"
        self._h = None
        self._dims = None
        self._staged = []

    def __del__(self):
        if getattr(self, "_h", None):
            lib.gcrl_replay_destroy(self._h)
            self._h = None

    def _stream(self):
        return _lib.current_stream(self.device_index)

    def _ensure(self, D, A):
        if self._h is None:
            h = vp()
            check(lib.gcrl_replay_create(C.byref(h), self.device_index, self.max_len, D, A,
                                         1 if self._PRIORITIZED else 0, float(self.alpha)))
            self._h, self._dims = h, (D, A)
        elif self._dims != (D, A):
            raise ValueError(f"transition shape changed: {self._dims} -> {(D, A)}")

    @property
    def handle(self):
        self._flush()
        return self._h

    def push(self, state, action, reward, next_state, done):       # src/buffer.py:13-14 / :46-48
        s, a, ns = _to_np(state).reshape(-1), _to_np(action).reshape(-1), _to_np(next_state).reshape(-1)
        self._ensure(s.shape[0], a.shape[0])
        self._staged.append(np.concatenate([s, a, np.float32(reward).reshape(1), ns,
                                            np.float32(done).reshape(1)]).astype(np.float32, copy=False))
        if len(self._staged) >= self._FLUSH:
            self._flush()

    def push_rows(self, s, a, r, ns, d):
        """Vectorised push of n transitions (arrays [n, D], [n, A], [n], [n, D], [n])."""
        s, a, ns = (np.asarray(x, np.float32) for x in (s, a, ns))
        self._ensure(s.shape[1], a.shape[1])
        self._flush()
        rows = np.concatenate([s, a, np.asarray(r, np.float32).reshape(-1, 1), ns,
                               np.asarray(d, np.float32).reshape(-1, 1)], axis=1)
        rows = np.ascontiguousarray(rows, np.float32)
        check(lib.gcrl_replay_push(self._h, rows.shape[0], np_ptr(rows), self._stream()))

    def _flush(self):
        if self._staged:
            rows = np.ascontiguousarray(np.stack(self._staged), np.float32)
            self._staged = []
            check(lib.gcrl_replay_push(self._h, rows.shape[0], np_ptr(rows), self._stream()))

    def __len__(self):
        committed = int(lib.gcrl_replay_len(self._h)) if self._h else 0
        return min(self.max_len, committed + len(self._staged))

    def _outputs(self, B):
        import torch
        D, A = self._dims
        dev = torch.device("cuda", self.device_index)
        return [torch.empty((B, w), dtype=torch.float32, device=dev) for w in (D, A, 1, D, 1)]

    def rows(self, first=0, n=None):
        """Stored transitions [first, first + n) in deque order as packed host rows (tests, checkpoints)."""
        self._flush()
        n = len(self) - first if n is None else n
        D, A = self._dims
        out = np.empty((n, 2 * D + A + 2), np.float32)
        check(lib.gcrl_replay_get_rows(self._h, first, n, np_ptr(out), self._stream()))
        return out

    # -- true resume: the live window in deque order (and the priorities), cf. HERBuffer.state_dict ----------
    def state_dict(self):
        self._flush()
        sd = {"kind": type(self).__name__, "max_len": self.max_len, "alpha": self.alpha, "dims": self._dims,
              "rows": self.rows() if self._h else None}
        if self._PRIORITIZED:
            sd["priorities"] = self.priorities if self._h else None
        return sd

    def load_state_dict(self, sd):
        if sd["kind"] != type(self).__name__ or sd["max_len"] != self.max_len:
            raise ValueError("buffer checkpoint was written by a different buffer type / max_len")
        self._staged = []
        if self._h:
            check(lib.gcrl_replay_destroy(self._h))
            self._h, self._dims = None, None
        if sd["rows"] is None or len(sd["rows"]) == 0:
            return
        self._ensure(*sd["dims"])
        rows = np.ascontiguousarray(sd["rows"], np.float32)
        check(lib.gcrl_replay_push(self._h, rows.shape[0], np_ptr(rows), self._stream()))
        if self._PRIORITIZED:
            self.set_priorities(sd["priorities"])


class ReplayBuffer(_Ring):
    """Reference src/buffer.py:8-35.  Positions come from ``random.sample(range(len), B)`` -- the same
    Mersenne-Twister draws as the reference's ``random.sample(self.buffer, B)`` -- through the C mirror."""

    def __init__(self, max_len: int, *, device=0):
        super().__init__(max_len, device=device)

    def sample_into(self, batch_size, out, indices=None):
        assert len(self) >= batch_size, "Not enough in buffer to sample"
        self._flush()
        B = int(batch_size)
        if indices is None:
            indices = _lib.py_sample_range(len(self), B)
        indices = np.ascontiguousarray(indices, np.int64)
        check(lib.gcrl_replay_sample(self._h, B, np_ptr(indices), *(vp(t.data_ptr()) for t in out), self._stream()))

    def sample(self, batch_size: int, indices=None):                            # :16-32
        assert len(self) >= batch_size, "Not enough in buffer to sample"
        self._flush()
        out = self._outputs(int(batch_size))
        self.sample_into(batch_size, out, indices)
        return tuple(out)


class PERBuffer(_Ring):
    """Reference src/buffer.py:38-89.  ``sample`` consumes ``np.random.random_sample(B)`` from NumPy's global
    legacy stream -- exactly what the reference's ``np.random.choice(N, B, p=P)`` draws -- and the positions
    are bit-identical to the reference's given the same priorities."""

    _PRIORITIZED = True

    def __init__(self, max_len: int, alpha: float, *, device=0):
        super().__init__(max_len, alpha, device=device)
        self.epsilon = 1e-6
        self._weights = None

    def sample_into(self, batch_size, beta, out, weights_ptr, uniforms=None, want_indices=False):
        """The device-side sample: batch -> ``out`` tensors, weights -> ``weights_ptr``; asynchronous unless
        the drawn positions are wanted on the host."""
        assert len(self) >= batch_size, "Not enough in buffer to sample"
        self._flush()
        B = int(batch_size)
        u = np.random.random_sample(B) if uniforms is None else uniforms
        u = np.ascontiguousarray(u, np.float64)
        idx = np.empty(B, np.int64) if want_indices else None
        check(lib.gcrl_replay_sample_prioritized(self._h, B, np_ptr(u), float(beta), *(vp(t.data_ptr()) for t in out),
                                                 weights_ptr, np_ptr(idx) if idx is not None else None,
                                                 self._stream()))
        return idx

    def sample(self, batch_size: int, beta: float, uniforms=None):              # :50-81
        import torch
        B = int(batch_size)
        assert len(self) >= B, "Not enough in buffer to sample"
        self._flush()
        out = self._outputs(B)
        w = torch.empty((B, 1), dtype=torch.float32, device=out[0].device)
        idx = self.sample_into(B, beta, out, vp(w.data_ptr()), uniforms, want_indices=True)
        return (*out, w, idx)

    def update_priorities(self, indices, priorities):                           # :86-89
        import torch
        B = len(indices)
        idx = np.ascontiguousarray(indices, np.int64)
        if isinstance(priorities, torch.Tensor):
            td = priorities.to(device=self.device, dtype=torch.float32).reshape(-1).contiguous()
        else:
            td = torch.from_numpy(np.ascontiguousarray(np.asarray(priorities, np.float32).reshape(-1))).to(self.device)
        check(lib.gcrl_replay_update_priorities(self._h, B, np_ptr(idx), vp(td.data_ptr()), self._stream()))
        torch.cuda.current_stream(self.device_index).synchronize()     # td may be freed by the caller

    def update_priorities_last(self, batch_size, td_ptr):
        """update_priorities for the positions of the preceding sample, TD errors read from device memory."""
        check(lib.gcrl_replay_update_priorities(self._h, int(batch_size), None, td_ptr, self._stream()))

    def last_positions(self, batch_size):
        """The deque positions the preceding sample drew (the reference's ``indices``)."""
        idx = np.empty(int(batch_size), np.int64)
        check(lib.gcrl_replay_last_positions(self._h, int(batch_size), np_ptr(idx), self._stream()))
        return idx

    @property
    def priorities(self):
        """The priorities in deque order (host float32 copy)."""
        self._flush()
        out = np.empty(len(self), np.float32)
        if len(self):
            check(lib.gcrl_replay_get_priorities(self._h, np_ptr(out), self._stream()))
        return out

    def set_priorities(self, prio):
        self._flush()
        prio = np.ascontiguousarray(prio, np.float32)
        check(lib.gcrl_replay_set_priorities(self._h, np_ptr(prio), prio.shape[0], self._stream()))

    def last_sample_info(self):
        """(float32 priority sum, True if the float64 cumsum took the sequential fallback) of the last sample."""
        s, f = C.c_float(), C.c_int()
        check(lib.gcrl_replay_last_sample_info(self._h, C.byref(s), C.byref(f), self._stream()))
        return np.float32(s.value), bool(f.value)

    def last_tables(self):
        """(P float32 [len], searched float64 table [len]) of the last sample."""
        n = len(self)
        p, cdf = np.empty(n, np.float32), np.empty(n, np.float64)
        check(lib.gcrl_replay_last_tables(self._h, np_ptr(p), np_ptr(cdf), self._stream()))
        return p, cdf
