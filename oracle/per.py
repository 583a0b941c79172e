"""Oracle: the two non-HER buffers behind ``agent.update`` -- uniform replay and prioritised replay.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

Restates (NumPy / pure Python, CPU):
  * ``ReplayBuffer``   src/buffer.py:8-35   (``random.sample`` over a bounded deque)
  * ``PERBuffer``      src/buffer.py:38-89  (priority deque, ``np.random.choice(N, B, p=P)``,
                                             importance weights, ``update_priorities``)
and the third-party arithmetic the PER path goes through, written out so that the CUDA kernels have an
explicit rule to match bit for bit:
  * NumPy's float32 ``ndarray.sum()`` on a contiguous vector: pairwise summation, 8 strided accumulators
    on leaves of at most 128 elements, halves rounded down to a multiple of 8, seeded with the additive
    identity (``pairwise_sum_f32``);
  * the legacy ``RandomState.choice`` with replacement and ``p``: ``cdf = p.astype(float64).cumsum()``
    (sequential left-to-right float64 additions), ``cdf /= cdf[-1]``, ``u = random_sample(B)``,
    ``searchsorted(cdf, u, side="right")`` (``choice_indices``).
Both restatements are pinned against NumPy itself in tests/test_per_oracle_cpu.py, and the classes against
the unmodified reference classes through tests/golden (make_golden.py ``per``).

Float32 ``power`` (importance weights ``(N p_i)^-beta`` and priorities ``(|td| + 1e-6)^alpha``) goes through
the platform's ``powf`` in the reference (libm or a SIMD library, depending on the CPU) and is therefore only
reproducible to an ulp; the parity tests hold those two quantities to 2 ulp of float32 and everything that
is integer or index work (positions, FIFO eviction, the searchsorted draws given the priorities) exactly.
"""
from __future__ import annotations

import random
from collections import deque

import numpy as np

F32 = np.float32


def pairwise_sum_f32(a) -> np.float32:
    """``np.asarray(a, float32).sum()`` restated (NumPy ``FLOAT_pairwise_sum`` + identity-seeded reduce)."""
    a = np.ascontiguousarray(a, dtype=F32)

    def rec(lo, n):
        if n < 8:
            r = F32(0.0)
            for i in range(n):
                r = F32(r + a[lo + i])
            return r
        if n <= 128:
            r = a[lo:lo + 8].copy()
            end = n - (n % 8)
            for i in range(8, end, 8):
                r += a[lo + i:lo + i + 8]
            res = F32(F32(F32(r[0] + r[1]) + F32(r[2] + r[3])) + F32(F32(r[4] + r[5]) + F32(r[6] + r[7])))
            for i in range(end, n):
                res = F32(res + a[lo + i])
            return res
        n2 = n // 2
        n2 -= n2 % 8
        return F32(rec(lo, n2) + rec(lo + n2, n - n2))

    return F32(F32(0.0) + rec(0, a.shape[0]))


def pairwise_leaves(n):
    """The leaf segments (offset, length) of the summation tree above, left to right, and the internal nodes
    as (left, right) child ids in an order in which children precede parents (leaves are ids 0..L-1, internal
    node j is id L+j; the last one is the root).  Used by tests to check the device-side tree."""
    leaves, internal = [], []

    def rec(lo, m):
        if m <= 128:
            leaves.append((lo, m))
            return -len(leaves)                 # leaves numbered -1, -2, ... until L is known
        m2 = m // 2
        m2 -= m2 % 8
        l = rec(lo, m2)
        r = rec(lo + m2, m - m2)
        internal.append((l, r))
        return len(internal) - 1

    rec(0, n)
    L = len(leaves)
    fix = lambda c: (-c - 1) if c < 0 else L + c          # noqa: E731
    return leaves, [(fix(l), fix(r)) for l, r in internal]


def normalised_priorities(prio) -> np.ndarray:
    """src/buffer.py:53-59: float32 vector, divided by its float32 sum (uniform when the sum is not > 0)."""
    P = np.array(prio, dtype=F32)
    s = pairwise_sum_f32(P)
    if s > 0:
        P /= s
    else:
        P[:] = 1.0 / P.shape[0]
    return P


def choice_cdf(P) -> np.ndarray:
    """The float64 table ``RandomState.choice`` searches: sequential cumsum, divided by its last entry."""
    p = np.asarray(P, dtype=np.float64)
    cdf = np.empty_like(p)
    acc = 0.0
    for i in range(p.shape[0]):                 # np.cumsum is a plain left-to-right accumulate
        acc = acc + float(p[i])
        cdf[i] = acc
    cdf /= cdf[-1]
    return cdf


def choice_indices(P, u) -> np.ndarray:
    """``np.random.choice(N, B, p=P)`` given the B uniforms ``u = random_sample(B)`` it draws."""
    return np.searchsorted(choice_cdf(P), np.asarray(u, np.float64), side="right").astype(np.int64)


def importance_weights(P, indices, beta) -> np.ndarray:
    """src/buffer.py:65-66."""
    N = P.shape[0]
    w = (N * P[indices]) ** (-beta)
    w /= w.max()
    return w.astype(F32)


def new_priority(td, alpha, epsilon=1e-6) -> np.float32:
    """src/buffer.py:89 for one float32 TD error."""
    return F32((abs(F32(td)) + epsilon) ** alpha)


class ReplayBufferOracle:
    """src/buffer.py:8-35 with float32 NumPy rows for states (the reference holds torch tensors)."""

    def __init__(self, max_len: int, sample=None):
        self.buffer = deque(maxlen=max_len)
        self._sample = sample or (lambda n, k: random.sample(range(n), k))
        self.last_indices = None

    def push(self, state, action, reward, next_state, done):
        self.buffer.append((np.asarray(state, F32), np.asarray(action, F32), F32(reward),
                            np.asarray(next_state, F32), F32(done)))

    def sample(self, batch_size: int):
        assert len(self.buffer) >= batch_size, "Not enough in buffer to sample"
        idx = self._sample(len(self.buffer), batch_size)      # same stream as random.sample(deque, k)
        self.last_indices = np.asarray(idx, np.int64)
        return _collate([self.buffer[i] for i in idx])

    def __len__(self):
        return len(self.buffer)


class PERBufferOracle:
    """src/buffer.py:38-89.  ``uniforms(B)`` defaults to NumPy's global legacy stream, which is what
    ``np.random.choice`` consumes."""

    def __init__(self, max_len: int, alpha: float, uniforms=None):
        self.buffer = deque(maxlen=max_len)
        self.priorities = deque(maxlen=max_len)
        self.alpha = alpha
        self.epsilon = 1e-6
        self._uniforms = uniforms or (lambda b: np.random.random_sample(b))

    def push(self, state, action, reward, next_state, done):
        self.buffer.append((np.asarray(state, F32), np.asarray(action, F32), F32(reward),
                            np.asarray(next_state, F32), F32(done)))
        self.priorities.append(1.0)

    def sample(self, batch_size: int, beta: float):
        assert len(self) >= batch_size, "Not enough in buffer to sample"
        P = normalised_priorities(self.priorities)
        indices = choice_indices(P, self._uniforms(batch_size))
        s, a, r, ns, d = _collate([self.buffer[i] for i in indices])
        w = importance_weights(P, indices, beta)
        return s, a, r, ns, d, w[:, None], indices

    def __len__(self):
        return len(self.buffer)

    def update_priorities(self, indices, priorities):
        priorities = np.asarray(priorities, F32).reshape(-1)
        for index, priority in zip(indices, priorities):       # in order: the last duplicate wins
            self.priorities[index] = new_priority(priority, self.alpha, self.epsilon)


def _collate(rows):
    s, a, r, ns, d = zip(*rows)
    return (np.stack(s).astype(F32), np.stack(a).astype(F32), np.asarray(r, F32)[:, None],
            np.stack(ns).astype(F32), np.asarray(d, F32)[:, None])
