import sys, os
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "goal-conditioned-rl-framework_b200"))
import torch
from gcrl_b200._lib import lib, check, vp
N = K = 256
for M in (1024, 2048, 4096, 8192, 16384, 65536):
    x = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda") / 16; b = torch.randn(N, device="cuda")
    y = torch.empty(M, N, device="cuda"); st = vp(torch.cuda.current_stream().cuda_stream)
    res = []
    for engine in (0, 1):
        def go(n):
            for _ in range(n):
                check(lib.gcrl_dense_layer(0, engine, 0, M, N, K, vp(x.data_ptr()), K, vp(w.data_ptr()), K, vp(b.data_ptr()), None, 0, vp(y.data_ptr()), N, st))
        go(3)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); go(20); e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 20 * 1000)
    print(f"BN={os.environ.get('GCRL_TC_BN','auto')} M={M}: ffma {res[0]:.1f} us  tc {res[1]:.1f} us")
