#!/usr/bin/env python
"""CUDA-event timing of the HER sampler alone (L2 flushed between launches) on bench.py's 1M-transition buffer.
Usage: python profiles/time_sampler.py [B ...]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "goal-conditioned-rl-framework_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from gcrl_b200 import HERBuffer  # noqa: E402
from gcrl_b200._lib import check, lib, vp  # noqa: E402

batches = [int(x) for x in sys.argv[1:]] or [256, 1024, 4096, 16384, 65536, 1 << 20]
sys.argv = sys.argv[:1]
args = bench.parse()
T, k, O, G, A = 50, args.k_future, args.obs, args.goal, args.act
D = O + G
E = int(os.environ.get("PROF_EPISODES", "20000"))
data = bench.synth(np.random.default_rng(0), E, T, O, G, A, k)
buf = HERBuffer(E * 246, 50, 1, k_future=k, index_source="device", seed=7)
for e in range(E):
    buf.push_episode(data["s"][e], data["a"][e], data["ns"][e], data["r"][e], data["d"][e], data["ag"][e], data["fut"][e])
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
stream = torch.cuda.current_stream(dev)
sp = vp(stream.cuda_stream)
alg = 4 * (2 * O + A + 2 * G) + 5 + 4 * (2 * D + A + 2)
for B in batches:
    outs = [torch.empty((B, w), dtype=torch.float32, device=dev) for w in (D, A, 1, D, 1)]
    ptrs = [vp(o.data_ptr()) for o in outs]
    for _ in range(5):
        check(lib.gcrl_her_sample(buf.handle, B, None, *ptrs, None, sp))
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(30)]
    for i, (a0, a1) in enumerate(evs):
        flush.fill_(i & 0xFF)
        a0.record(stream)
        check(lib.gcrl_her_sample(buf.handle, B, None, *ptrs, None, sp))
        a1.record(stream)
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in evs) / len(evs)
    print(f"B={B:8d}: {ms * 1e3:8.2f} us  {B * alg / (ms * 1e-3) / 1e9:8.1f} GB/s algorithmic  ({B / (ms * 1e-3) / 1e6:.1f} M transitions/s)")
