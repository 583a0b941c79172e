"""Oracle: the product's on-device index stream, restated in NumPy.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

The reference draws sample positions with ``random.sample`` (src/buffer.py:124); the
product offers that exact stream from the host ("host" index source) and, for
throughput, an on-device stream: position i of call ``epoch`` is a keyed 4-round
Feistel permutation of [0, n) with cycle walking (csrc/her.cu ``feistel_position``),
i.e. B distinct uniformly spread positions -- sampling without replacement, like
``random.sample``.  This restatement lets the GPU tests check the device stream
bit-for-bit and the gathered rows against the eager oracle at those positions.
"""
import numpy as np

M64 = (1 << 64) - 1


def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M64
    return x ^ (x >> 31)


def mix32(x):
    x = np.asarray(x, np.uint64)
    x = x ^ (x >> np.uint64(16))
    x = (x * np.uint64(0x7FEB352D)) & np.uint64(0xFFFFFFFF)
    x = x ^ (x >> np.uint64(15))
    x = (x * np.uint64(0x846CA68B)) & np.uint64(0xFFFFFFFF)
    x = x ^ (x >> np.uint64(16))
    return x


def feistel_positions(i, n, seed, epoch):
    i = np.asarray(i, np.uint64).copy()
    if n <= 1:
        return np.zeros(i.shape, np.int64)
    b = 0
    while b < 63 and (1 << b) < n:
        b += 1
    b = max(b, 2)
    b += b & 1
    half = np.uint64(b >> 1)
    mask = np.uint64((1 << (b >> 1)) - 1)
    k0 = splitmix64(seed ^ ((epoch * 0xD1B54A32D192ED03) & M64))
    k1 = splitmix64(k0)
    keys = [np.uint64(k0 & 0xFFFFFFFF), np.uint64(k0 >> 32), np.uint64(k1 & 0xFFFFFFFF),
            np.uint64(k1 >> 32)]
    x = i
    todo = np.ones(x.shape, bool)
    while todo.any():
        xs = x[todo]
        L = (xs >> half) & mask
        R = xs & mask
        for r in range(4):
            L, R = R, L ^ (mix32(R ^ keys[r]) & mask)
        xs = (L << half) | R
        x[todo] = xs
        todo[todo] = xs >= np.uint64(n)
    return x.astype(np.int64)
