"""Time gcrl_dense_layer_presplit at M = 65536, N = K = 256 (one hidden layer of the B = 65536 update) with CUDA events,
L2 flushed between launches; the environment selects the kernel variant (GCRL_TC_PAIR, GCRL_TC_DBG)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "goal-conditioned-rl-framework_b200"))
from gcrl_b200._lib import check, lib, vp  # noqa: E402

M, N, K = int(os.environ.get("TC_M", 65536)), 256, int(os.environ.get("TC_K", 256))
torch.manual_seed(0)
x = torch.randn(M, K, device="cuda")
w = torch.randn(N, K, device="cuda") / 16
b = torch.randn(N, device="cuda")
y = torch.empty(M, N, device="cuda")
hi, lo = torch.empty_like(w), torch.empty_like(w)
st = vp(torch.cuda.current_stream().cuda_stream)
check(lib.gcrl_split_tf32(0, vp(w.data_ptr()), vp(hi.data_ptr()), vp(lo.data_ptr()), w.numel(), st))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def run():
    check(lib.gcrl_dense_layer_presplit(0, 0, M, N, K, vp(x.data_ptr()), K, vp(hi.data_ptr()), vp(lo.data_ptr()), K,
                                        vp(b.data_ptr()), None, 0, vp(y.data_ptr()), N, st))


FLUSH = os.environ.get("TC_FLUSH", "write")


def timed(pair, dbg):
    os.environ["GCRL_TC_PAIR_RT"] = str(pair)
    os.environ["GCRL_TC_DBG"] = str(dbg)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(15):
        if FLUSH == "write":
            flush.zero_()
        elif FLUSH == "read":
            sink = flush.sum(dtype=torch.int64)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print(f"flush={FLUSH} pair={pair} dbg={dbg:3d} M={M} K={K}: median {ts[len(ts) // 2]:.1f} us, min {ts[0]:.1f} us", flush=True)


for pair in [int(v) for v in os.environ.get("TC_PAIRS", "1,0").split(",")]:
    for dbg in [int(v) for v in os.environ.get("TC_DBGS", "0,32,3,35,2,34").split(",")]:
        timed(pair, dbg)
