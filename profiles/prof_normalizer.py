#!/usr/bin/env python
"""Driver for ncu captures of the normaliser kernels: RunningNormalizer.update / normalize on a device-resident
float32 batch [n, 19].  Usage: python profiles/prof_normalizer.py [n]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "goal-conditioned-rl-framework_b200"))
import torch  # noqa: E402

from gcrl_b200 import RunningNormalizer  # noqa: E402
from gcrl_b200._lib import check, lib, vp  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dim = 19
dev = torch.device("cuda", 0)
sp = vp(torch.cuda.current_stream(dev).cuda_stream)
nz = RunningNormalizer(dim, device=0)
x = torch.randn(n, dim, device=dev)
y = torch.empty(n, dim, device=dev)
for _ in range(3):
    check(lib.gcrl_norm_update_dev(nz._h, vp(x.data_ptr()), n, 0, sp))
    check(lib.gcrl_norm_apply_dev_f32(nz._h, vp(x.data_ptr()), n, 0, vp(y.data_ptr()), dim, 0, sp))
torch.cuda.synchronize()
print("ok", n, float(y.abs().max()))
