import sys, os, ctypes as C, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "goal-conditioned-rl-framework_b200"))
import numpy as np, torch
from gcrl_b200._lib import lib, check, vp
def run(engine, mode, x, w, b, act, N, K):
    M = x.shape[0]
    y = torch.full((M, w.shape[0]), float("nan"), device="cuda")
    st = vp(torch.cuda.current_stream().cuda_stream)
    check(lib.gcrl_dense_layer(0, engine, mode, M, N, K, vp(x.data_ptr()), x.stride(0), vp(w.data_ptr()), w.stride(0),
          vp(b.data_ptr()) if b is not None else None, vp(act.data_ptr()) if act is not None else None,
          act.stride(0) if act is not None else 0, vp(y.data_ptr()), y.stride(0), st))
    torch.cuda.synchronize()
    return y
torch.manual_seed(0)
for (M, N, K) in [(128, 256, 256), (1000, 256, 256), (65536, 256, 256), (300, 64, 64), (4096, 512, 512), (777, 96, 100), (5000, 256, 24)]:
    x = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda") / K ** 0.5; b = torch.randn(N, device="cuda")
    act = torch.randn(M, N, device="cuda")
    ref = (x.double() @ w.double().T)
    for mode in (0, 1, 2):
        if mode == 0: want = torch.nn.functional.leaky_relu(ref + b.double(), 0.01)
        elif mode == 1: want = ref * torch.where(act > 0, 1.0, 0.01).double()
        else: want = ref + b.double()
        got = run(1, mode, x, w, b, act, N, K)
        err = ((got.double() - want).abs().max() / want.abs().max()).item()
        msg = f"M={M} N={N} K={K} mode={mode} tc rel err {err:.3e}"
        if mode != 1:
            g0 = run(0, mode, x, w, b, act, N, K)
            msg += f"  ffma rel err {((g0.double() - want).abs().max() / want.abs().max()).item():.3e}"
        print(msg, flush=True)
# timing
M, N, K = 65536, 256, 256
x = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda") / 16; b = torch.randn(N, device="cuda")
for engine in (0, 1):
    for _ in range(3): run(engine, 0, x, w, b, None, N, K)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    y = torch.empty(M, N, device="cuda"); st = vp(torch.cuda.current_stream().cuda_stream)
    e0.record()
    for _ in range(20):
        check(lib.gcrl_dense_layer(0, engine, 0, M, N, K, vp(x.data_ptr()), K, vp(w.data_ptr()), K, vp(b.data_ptr()), None, 0, vp(y.data_ptr()), N, st))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"engine {engine}: {ms*1000:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s (fp32-equivalent)")
