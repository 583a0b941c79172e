"""Oracle: HER replay buffer, sparse Panda reward, running normaliser.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

Restates (NumPy / pure Python, CPU):
  * ``panda_gym`` sparse ``compute_reward`` (bound at src/env.py:105, called at
    src/buffer.py:166).  PARITY UNPINNED: third-party, un-vendored, un-pinned.
  * ``HERBuffer``            src/buffer.py:92-179
  * ``RunningNormalizer``    src/utils.py:68-117

The buffer port keeps the reference's *eager* structure on purpose (relabelled
copies are materialised into a bounded deque at episode end; ``sample`` is a
uniform draw without replacement over that deque), because it is also the timed
CPU baseline in ``bench.py`` and must cost what the reference costs.  The CUDA
product is lazy (relabels at sample time); the parity tests prove the two are
entry-for-entry identical on the same RNG draws.
"""
from __future__ import annotations

import random
from collections import deque

import numpy as np

DISTANCE_THRESHOLD = np.float32(0.05)


def compute_reward(achieved_goal, desired_goal, info=None):
    """Sparse Panda reward: ``-(||a - d||_2 > 0.05)`` as float32 (-1.0 or -0.0).

    Follows the call site src/buffer.py:166 ``compute_reward(ag, future_ag, {})``.
    Pinned arithmetic rule (the one the CUDA kernel implements bit-for-bit):
    float32 inputs, squared differences summed left to right in float32 without
    fused multiply-add, float32 square root, compared against float32(0.05);
    success yields -0.0 (sign bit set), exactly what ``-np.array(False, float32)``
    gives.
    """
    a = np.asarray(achieved_goal, dtype=np.float32)
    d = np.asarray(desired_goal, dtype=np.float32)
    diff = a - d
    sq = diff * diff
    acc = sq[..., 0]
    for i in range(1, sq.shape[-1]):
        acc = (acc + sq[..., i]).astype(np.float32)
    dist = np.sqrt(acc, dtype=np.float32)
    return -np.array(dist > DISTANCE_THRESHOLD, dtype=np.float32)


class HERBufferOracle:
    """Eager future-strategy HER buffer (src/buffer.py:92-179).

    ``randint`` / ``sample`` default to Python's global Mersenne-Twister, which
    is the stream the reference consumes (src/buffer.py:124,153); tests may pass
    recording wrappers.  States are plain float32 NumPy rows here (the reference
    holds torch tensors and converts them at src/buffer.py:147-148).
    """

    FLUSH_LEN = 50  # literal at src/buffer.py:117 (not max_eps_len)

    def __init__(self, max_mem_len, max_eps_len, nenvs, threshold=0.05, k_future=4,
                 randint=None, sample=None, reward_fn=compute_reward):
        self.buffer = deque(maxlen=max_mem_len)                      # :101
        self.episodes = [deque(maxlen=max_eps_len) for _ in range(nenvs)]  # :102
        self.threshold = threshold
        self.k_future = k_future
        self.compute_reward = reward_fn
        self._randint = randint or random.randint
        self._sample = sample or random.sample

    def __len__(self):                                               # :137-138
        return len(self.buffer)

    def push(self, idx, state, action, next_state, reward, done, desired_goal,
             achieved_goal):                                         # :110-119
        self.episodes[idx].append(
            (np.asarray(state, np.float32), action, np.asarray(next_state, np.float32),
             reward, done, desired_goal, achieved_goal))
        if done or len(self.episodes[idx]) >= self.FLUSH_LEN:
            self.apply_her(idx)
            self.episodes[idx].clear()

    def apply_her(self, idx):                                        # :143-179
        ep = self.episodes[idx]
        T = len(ep)
        for t, (s, a, ns, r, d, dg, ag) in enumerate(ep):
            self.buffer.append((s, a, ns, r, d, dg, ag))             # original, :149
            for _ in range(self.k_future):                           # :151
                if t >= T - 1:                                       # :152
                    continue
                f = self._randint(t + 1, T - 1)                      # inclusive, :153
                future_ag = ep[f][6]                                 # :154
                new_goal = np.array(future_ag, dtype=np.float32)     # :156
                G = new_goal.shape[0]
                s2 = np.concatenate([s[:-G], new_goal], axis=-1)     # :159-161
                ns2 = np.concatenate([ns[:-G], new_goal], axis=-1)   # :162-164
                r2 = self.compute_reward(ag, future_ag, {})          # :166
                self.buffer.append((s2, a, ns2, r2, False, new_goal, ag))  # :169-179

    def sample(self, batch_size):                                    # :121-135
        assert len(self.buffer) >= batch_size, "[ERROR] Not enough in buffer to sample"
        rows = self._sample(self.buffer, batch_size)                 # :124
        return collate(rows)


def collate(rows):
    """5 float32 arrays exactly as src/buffer.py:125-133 builds them."""
    s, a, ns, r, d, _, _ = zip(*rows)
    states = np.array(s).astype(np.float32)
    actions = np.array(a).astype(np.float32)
    rewards = np.array([np.float32(x) for x in r], dtype=np.float32)[:, None]
    next_states = np.array(ns).astype(np.float32)
    dones = np.array([np.float32(bool(x)) for x in d], dtype=np.float32)[:, None]
    return states, actions, rewards, next_states, dones


def materialise_episode(s, a, ns, r, d, ag, fut, k_future):
    """Vectorised eager expansion of ONE episode into its stored entries.

    Same entry order as ``apply_her`` (src/buffer.py:145-179): for each t the
    original, then (t < T-1) the k relabelled copies with future index
    ``fut[t, j]``.  Used for mid-size parity cases where the deque port is slow.
    Returns (states, actions, rewards, next_states, dones), each [E, ...] with
    E = (T-1)(k+1)+1.
    """
    T = s.shape[0]
    G = ag.shape[1]
    k = k_future
    reps = np.full(T, k + 1, dtype=np.int64)
    reps[T - 1] = 1
    tt = np.repeat(np.arange(T), reps)
    starts = np.cumsum(reps) - reps
    jj = np.arange(tt.shape[0]) - starts[tt]
    S = s[tt].astype(np.float32).copy()
    NS = ns[tt].astype(np.float32).copy()
    A = a[tt].astype(np.float32)
    R = r[tt].astype(np.float32).copy()
    Dn = d[tt].astype(np.float32).copy()
    rel = jj > 0
    if rel.any():
        f = fut[tt[rel], jj[rel] - 1].astype(np.int64)
        g = ag[f].astype(np.float32)
        S[rel, -G:] = g
        NS[rel, -G:] = g
        R[rel] = compute_reward(ag[tt[rel]], ag[f])
        Dn[rel] = 0.0
    return S, A, R[:, None], NS, Dn[:, None]


class RunningNormalizerOracle:
    """src/utils.py:68-117 (Chan parallel-variance merge, float64 state)."""

    def __init__(self, size, clip_range=5.0, eps=1e-8):              # :69-73
        self.mean = np.zeros(size)
        self.var = np.ones(size)
        self.count = eps
        self.clip_range = clip_range

    def update(self, x):                                             # :75-80
        x = np.asarray(x)
        self._update_from_moments(np.mean(x, axis=0), np.var(x, axis=0), x.shape[0])

    def _update_from_moments(self, mean, var, count):                # :82-94
        total = self.count + count
        delta = mean - self.mean
        new_mean = self.mean + delta * count / total
        m2 = self.var * self.count + var * count \
            + np.square(delta) * self.count * count / total
        self.mean, self.var, self.count = new_mean, m2 / total, total

    def normalize(self, x):                                          # :96-98
        z = (x - self.mean) / (np.sqrt(self.var) + 1e-8)
        return np.clip(z, -self.clip_range, self.clip_range)
