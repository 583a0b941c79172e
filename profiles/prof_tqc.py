#!/usr/bin/env python
"""Driver for launch lists / ncu captures of the SAC / TQC update: BASELINE configs[3] shape (Slide = Push shape,
hidden 512 x 3 as config_tqc_push.yaml, batch 512).  Usage: python profiles/prof_tqc.py [tqc|sac] [B] [H] [updates]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "goal-conditioned-rl-framework_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from gcrl_b200 import SACAgent, TQCAgent  # noqa: E402

algo = sys.argv[1] if len(sys.argv) > 1 else "tqc"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
H = int(sys.argv[3]) if len(sys.argv) > 3 else 512
updates = int(sys.argv[4]) if len(sys.argv) > 4 else 8
sys.argv = sys.argv[:1]
args = bench.parse()
args.batch, args.hidden = B, H
T, k, O, G, A = 50, args.k_future, args.obs, args.goal, args.act
E = 400
data = bench.synth(np.random.default_rng(0), E, T, O, G, A, k)
cfg = bench.agent_config(args, E * 246)
cfg.alpha_lr, cfg.alpha_min, cfg.alpha_min_steps, cfg.grad_clip, cfg.gamma = 3e-4, 3e-4, 1, 5.0, 0.95
torch.manual_seed(0)
ag = (TQCAgent if algo == "tqc" else SACAgent)(O + G, A, cfg, None, 1, 40, index_source="device")
for e in range(E):
    ag.buffer.push_episode(data["s"][e], data["a"][e], data["ns"][e], data["r"][e], data["d"][e], data["ag"][e], data["fut"][e])
import time
for i in range(3):
    ag.update(i + 1)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(updates):
    info = ag.update(4 + i)
torch.cuda.synchronize()
print("ok", algo, B, H, f"{(time.perf_counter() - t0) / updates * 1e3:.3f} ms per update", [float(np.mean(x)) for x in info][:3])
