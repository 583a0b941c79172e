import sys, os, ctypes as C
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "goal-conditioned-rl-framework_b200"))
import torch
from gcrl_b200._lib import lib, check, vp
def run(engine, dz, x, N, K, ldw, max_splits=128):
    M = dz.shape[0]
    stride = N * ldw + 8
    pw = torch.full((max_splits, stride), float("nan"), device="cuda"); pb = torch.full((max_splits, N + 4), float("nan"), device="cuda")
    sp = C.c_int()
    check(lib.gcrl_dense_wgrad(0, engine, M, N, K, vp(dz.data_ptr()), dz.stride(0), vp(x.data_ptr()), x.stride(0), vp(pw.data_ptr()), ldw, stride,
          vp(pb.data_ptr()), N + 4, max_splits, C.byref(sp), vp(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    S = sp.value
    W = pw[:S, :N * ldw].double().sum(0).reshape(N, ldw)
    return W, pb[:S, :N].double().sum(0), S
torch.manual_seed(0)
for (M, N, K) in [(1024, 128, 32), (4096, 256, 256), (65536, 256, 256), (5000, 64, 64), (3000, 256, 24), (8192, 512, 512), (777, 96, 100), (2048, 1, 256), (2048, 4, 64)]:
    dz = torch.randn(M, N, device="cuda") / M ** 0.5; x = torch.randn(M, K, device="cuda")
    ldw = (K + 3) // 4 * 4
    want = dz.double().T @ x.double(); wb = dz.double().sum(0)
    for engine in (1, 0):
        try:
            W, b, S = run(engine, dz, x, N, K, ldw)
            err = ((W[:, :K] - want).abs().max() / want.abs().max()).item()
            errb = ((b - wb).abs().max() / wb.abs().max()).item()
            pad = float(W[:, K:].abs().max()) if ldw > K else 0.0
            print(f"M={M} N={N} K={K} engine={engine} slabs={S} rel err W {err:.2e} b {errb:.2e} pad {pad}", flush=True)
        except Exception as e:
            print(f"M={M} N={N} K={K} engine={engine} FAILED {e}", flush=True)
M, N, K = 65536, 256, 256
dz = torch.randn(M, N, device="cuda"); x = torch.randn(M, K, device="cuda")
for engine in (0, 1):
    for _ in range(3): run(engine, dz, x, N, K, K)
    pw = torch.empty((128, N * K + 8), device="cuda"); pb = torch.empty((128, N + 4), device="cuda"); sp = C.c_int()
    st = vp(torch.cuda.current_stream().cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        check(lib.gcrl_dense_wgrad(0, engine, M, N, K, vp(dz.data_ptr()), N, vp(x.data_ptr()), K, vp(pw.data_ptr()), K, N * K + 8, vp(pb.data_ptr()), N + 4, 128, C.byref(sp), st))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"wgrad engine {engine}: {ms*1000:.1f} us  {2*M*N*K/ms/1e9:.1f} TFLOP/s (fp32-equivalent), slabs {sp.value}")
