#!/usr/bin/env python
"""Driver for `ncu --set full` of the HER sampler alone: bench.py's 1M-transition buffer (20 000 episodes of 50
steps, k_future 4 -> 4.92M deque entries) and a few launches of her_sample_kernel at batch B (device index
stream).  Usage: python profiles/prof_sampler.py [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "goal-conditioned-rl-framework_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from gcrl_b200 import HERBuffer  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
sys.argv = sys.argv[:1]
args = bench.parse()
T, k, O, G, A = 50, args.k_future, args.obs, args.goal, args.act
E = int(os.environ.get("PROF_EPISODES", "20000"))
data = bench.synth(np.random.default_rng(0), E, T, O, G, A, k)
buf = HERBuffer(E * 246, 50, 1, k_future=k, index_source="device", seed=7)
for e in range(E):
    buf.push_episode(data["s"][e], data["a"][e], data["ns"][e], data["r"][e], data["d"][e], data["ag"][e], data["fut"][e])
torch.cuda.synchronize()
for _ in range(4):
    out = buf.sample(B)
torch.cuda.synchronize()
print("ok", B, len(buf), float(out[2].mean()))
