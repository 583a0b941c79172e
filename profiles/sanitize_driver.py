#!/usr/bin/env python
"""Small pass over every kernel family of the hot path, sized for compute-sanitizer (profiles/sanitize.sh)."""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "goal-conditioned-rl-framework_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from gcrl_b200 import DDPG, TD3Agent, RunningNormalizer  # noqa: E402


def cfg(B, H, buffer_type="HER", max_len=3000):
    return types.SimpleNamespace(
        hidden_dim=H, layer_count=3, actor_lr=1e-3, actor_lr_min=1e-3, ac_scheduler_steps=1, critic_lr=1e-3,
        critic_lr_min=1e-3, cr_scheduler_steps=1, buffer_type=buffer_type, max_len=max_len, alpha=0.6, batch_size=B,
        gamma=0.98, ac_update_freq=1, noise_std=0.2, noise_clamp=0.5, policy_noise=0.2, grad_clip=10.0, beta=0.4,
        beta_end=100, k_future=4, max_eps_len=50, tau=0.05)


rng = np.random.default_rng(0)
O, G, A, k = 18, 3, 3, 4
torch.manual_seed(0)
for cls, B, H in ((DDPG, 64, 64), (TD3Agent, 48, 64), (DDPG, 1100, 64), (DDPG, 2048, 64)):
    ag = cls(O + G, A, cfg(B, H), None, 1, 40, index_source="device", max_batch=B)
    for e in range(40):                                   # ragged episodes; the 3000-entry cap evicts the oldest
        T = 50 if e % 3 else int(rng.integers(1, 50))
        d = bench.synth(rng, 1, T, O, G, A, k)
        ag.buffer.push_episode(d["s"][0], d["a"][0], d["ns"][0], d["r"][0], d["d"][0], d["ag"][0], d["fut"][0])
    out = ag.buffer.sample(min(B, 256))
    for step in (39, 40, 41):
        info = ag.update(step)
    assert all(np.isfinite(float(x)) for x in info), info
    del ag
per = DDPG(O + G, A, cfg(64, 64, "PER", 4096), None, 1, 40)
d = bench.synth(rng, 60, 50, O, G, A, k)
per.buffer.push_rows(*(np.concatenate(list(d[key])) for key in ("s", "a", "r", "ns", "d")))
np.random.seed(0)
for step in (1, 2, 3):
    per.update(step)
nz = RunningNormalizer(19)
nz.update(rng.standard_normal((64, 19)))
nz.update(rng.standard_normal((20000, 19)).astype(np.float32))
q = nz.normalize(rng.standard_normal((300, 19)))
assert np.isfinite(q).all()
torch.cuda.synchronize()
print("sanitize driver OK")
