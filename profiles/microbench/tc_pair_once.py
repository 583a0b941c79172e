"""One launch of each dense-kernel variant (for ncu): pair kernel, then the single-CTA kernel."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "goal-conditioned-rl-framework_b200"))
from gcrl_b200._lib import check, lib, vp  # noqa: E402

M, N, K = 65536, 256, 256
x = torch.randn(M, K, device="cuda")
w = torch.randn(N, K, device="cuda") / 16
b = torch.randn(N, device="cuda")
y = torch.empty(M, N, device="cuda")
hi, lo = torch.empty_like(w), torch.empty_like(w)
st = vp(torch.cuda.current_stream().cuda_stream)
check(lib.gcrl_split_tf32(0, vp(w.data_ptr()), vp(hi.data_ptr()), vp(lo.data_ptr()), w.numel(), st))
for pair in (1, 0):
    os.environ["GCRL_TC_PAIR_RT"] = str(pair)
    for _ in range(2):
        check(lib.gcrl_dense_layer_presplit(0, 0, M, N, K, vp(x.data_ptr()), K, vp(hi.data_ptr()), vp(lo.data_ptr()), K,
                                            vp(b.data_ptr()), None, 0, vp(y.data_ptr()), N, st))
    torch.cuda.synchronize()
