#!/usr/bin/env python
"""Driver for the ncu launch list of a prioritised-replay update: DDPG.update(step) with buffer_type "PER" on N
stored transitions (bench.py's shapes).  The first updates run with all priorities 1.0 (no float64 addition of
the cumsum rounds: fixed-point scan); once TD errors have been written back the additions round and the
parity-function chunk kernels take over (PER_TDLIKE=1 starts there).  Usage: python profiles/prof_per.py [N] [updates]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "goal-conditioned-rl-framework_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from gcrl_b200 import DDPG  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
updates = int(sys.argv[2]) if len(sys.argv) > 2 else 12
sys.argv = sys.argv[:1]
args = bench.parse()
D, A, B = args.obs + args.goal, args.act, args.batch
cfg = bench.agent_config(args, N)
cfg.buffer_type, cfg.alpha, cfg.beta, cfg.beta_end = "PER", 0.6, 0.4, 10000
torch.manual_seed(0)
np.random.seed(0)
agent = DDPG(D, A, cfg, None, 1, 40, max_batch=B)
rng = np.random.default_rng(0)
for lo in range(0, N, 1 << 16):
    n = min(1 << 16, N - lo)
    s = rng.standard_normal((n, D)).astype(np.float32)
    agent.buffer.push_rows(s, rng.uniform(-1, 1, (n, A)).astype(np.float32), -(rng.random(n) > 0.3).astype(np.float32),
                           (s + 0.1 * rng.standard_normal((n, D))).astype(np.float32), np.zeros(n, np.float32))
if os.environ.get("PER_TDLIKE") == "1":      # priorities as after a long run: (|td| + 1e-6)^0.6, td ~ |N(0, 0.3)|
    agent.buffer.set_priorities(((np.abs(rng.normal(0, 0.3, N)) + 1e-6) ** 0.6).astype(np.float32))
torch.cuda.synchronize()
for step in range(1, updates + 1):
    info = agent.update(step)
torch.cuda.synchronize()
print("ok", N, B, float(info[0]), "cumsum additions round:", agent.buffer.last_sample_info()[1])
