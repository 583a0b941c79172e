import sys, os
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "goal-conditioned-rl-framework_b200"))
import torch
from gcrl_b200._lib import lib, check, vp
M, N, K = int(sys.argv[1]) if len(sys.argv) > 1 else 65536, 256, 256
x = torch.randn(M, K, device="cuda"); w = torch.randn(N, K, device="cuda") / 16; b = torch.randn(N, device="cuda")
y = torch.empty(M, N, device="cuda"); st = vp(torch.cuda.current_stream().cuda_stream)
for engine in (1, 1, 1, 0):
    check(lib.gcrl_dense_layer(0, engine, 0, M, N, K, vp(x.data_ptr()), K, vp(w.data_ptr()), K, vp(b.data_ptr()), None, 0, vp(y.data_ptr()), N, st))
torch.cuda.synchronize(); print("ok")
