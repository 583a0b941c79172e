#!/usr/bin/env python
"""Small driver for `ncu --set full` captures: one HER sample and a few DDPG updates at a given
batch, on a 2000-episode buffer (bench.py's workload with a short fill so that ncu's replay of
every launch stays cheap).  Usage: python profiles/prof_driver.py [B] [updates]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "goal-conditioned-rl-framework_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from gcrl_b200 import DDPG  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
updates = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sys.argv = sys.argv[:1]
args = bench.parse()
args.batch = B
T, k, O, G, A = 50, args.k_future, args.obs, args.goal, args.act
E = int(os.environ.get("PROF_EPISODES", "2000"))
data = bench.synth(np.random.default_rng(0), E, T, O, G, A, k)
torch.manual_seed(0)
agent = DDPG(O + G, A, bench.agent_config(args, E * 246), None, 1, 40, index_source="device", max_batch=B)
for e in range(E):
    agent.buffer.push_episode(data["s"][e], data["a"][e], data["ns"][e], data["r"][e], data["d"][e],
                              data["ag"][e], data["fut"][e])
torch.cuda.synchronize()
out = agent.buffer.sample(B)
for step in range(1, updates + 1):
    agent.update_async(step)
torch.cuda.synchronize()
print("ok", B, [float(x) for x in agent.read_metrics()][:4])
