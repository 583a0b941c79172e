#!/usr/bin/env python
"""Where the end-to-end microseconds of DDPG.update(step) go on the host (bench.py's e2e leg, piece by piece)."""
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "goal-conditioned-rl-framework_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from gcrl_b200 import DDPG  # noqa: E402
from gcrl_b200 import _lib  # noqa: E402
from gcrl_b200._lib import check, lib, np_ptr  # noqa: E402

sys.argv = sys.argv[:1]
args = bench.parse()
T, k, O, G, A, B = 50, args.k_future, args.obs, args.goal, args.act, 256
E = 20000
data = bench.synth(np.random.default_rng(0), E, T, O, G, A, k)
torch.manual_seed(0)
ag = DDPG(O + G, A, bench.agent_config(args, E * 246), None, 1, 40, index_source="host", max_batch=B)
for e in range(E):
    ag.buffer.push_episode(data["s"][e], data["a"][e], data["ns"][e], data["r"][e], data["d"][e], data["ag"][e], data["fut"][e])
random.seed(3)
for i in range(50):
    ag.update(i + 1)
torch.cuda.synchronize()


def timeit(name, fn, n=400):
    t0 = time.perf_counter()
    for i in range(n):
        fn(i)
    torch.cuda.synchronize()
    print(f"{name:58s} {(time.perf_counter() - t0) / n * 1e6:8.1f} us")


step = [100]


def full(i):
    ag.update(step[0]); step[0] += 1


timeit("agent.update(step) (host index stream, metrics read-back)", full)
idx = np.ascontiguousarray(np.random.default_rng(1).integers(0, len(ag.buffer), B), np.int64)


def given(i):
    ag.update(step[0], indices=idx); step[0] += 1


timeit("agent.update(step, indices=fixed) (no stream emulation)", given)
st = ag._stream()
m = ag._metrics
import ctypes as C
mp = C.cast(m, _lib.vp)
ip = np_ptr(idx)


def ccall(i):
    check(lib.gcrl_agent_update_from_buffer(ag._h, ag.buffer.handle, B, ip, None, 1e-3, 1e-3, 1, mp, st))


timeit("C call alone: update_from_buffer + metric poll", ccall)


def ccall_async(i):
    check(lib.gcrl_agent_update_from_buffer(ag._h, ag.buffer.handle, B, ip, None, 1e-3, 1e-3, 1, None, st))


timeit("C call alone, asynchronous (launch cost / GPU-bound rate)", ccall_async)
timeit("buffer.push_episode(..., None) (future offsets from the MT mirror)",
       lambda i: ag.buffer.push_episode(data["s"][i], data["a"][i], data["ns"][i], data["r"][i], data["d"][i], data["ag"][i], None))
timeit("_lib.py_sample_range(len, B) (the pre-draw)", lambda i: _lib.py_sample_range(len(ag.buffer), B))
