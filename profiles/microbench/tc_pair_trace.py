"""Timeline of the first CTA pair of tc_dense_pair_kernel (GCRL_TC_TRACE): per stage use, when each role handed over."""
import os
import sys

import numpy as np
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


def run():
    check(lib.gcrl_dense_layer_presplit(0, 0, M, N, K, vp(x.data_ptr()), K, vp(hi.data_ptr()), vp(lo.data_ptr()), K,
                                        vp(b.data_ptr()), None, 0, vp(y.data_ptr()), N, st))


for _ in range(3):
    run()
torch.cuda.synchronize()
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/tc_pair_trace.bin"
os.environ["GCRL_TC_TRACE"] = out
run()
torch.cuda.synchronize()
del os.environ["GCRL_TC_TRACE"]
t = np.fromfile(out, dtype=np.int64).reshape(2, 8, 64).astype(np.float64)
t0 = t[t > 0].min()
t = np.where(t > 0, t - t0, np.nan)
names = ["producer: empty seen", "splitter: full seen", "splitter: arrived", "MMA: stage ready", "MMA: committed",
         "relay: arrived", "epilogue: tfull seen", "epilogue: released"]
for cta in (0, 1):
    print(f"--- CTA {cta} (ns since first event; stage uses 0..23, then every 8th)")
    for r, nme in enumerate(names):
        row = t[cta, r]
        if np.all(np.isnan(row)):
            continue
        sel = list(range(24)) + list(range(24, 64, 8))
        print(f"{nme:24s}", " ".join("     -" if np.isnan(row[i]) else f"{row[i]:6.0f}" for i in sel))
