import sys, os, ctypes as C
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "goal-conditioned-rl-framework_b200"))
import torch
from gcrl_b200._lib import lib, check, vp
def run(dz, x, N, K):
    M = dz.shape[0]; ldw = K; stride = N * ldw
    pw = torch.full((128, stride), float("nan"), device="cuda"); sp = C.c_int()
    check(lib.gcrl_dense_wgrad(0, 1, M, N, K, vp(dz.data_ptr()), dz.stride(0), vp(x.data_ptr()), x.stride(0), vp(pw.data_ptr()), ldw, stride,
          None, 0, 1, C.byref(sp), vp(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    return pw[0].reshape(N, ldw), sp.value
M, N, K = 16, 128, 64
for (m0, n0, k0) in [(0, 0, 0), (0, 1, 0), (0, 0, 1), (0, 5, 9), (1, 0, 0), (3, 40, 33), (9, 100, 63), (0, 32, 0), (0, 0, 32), (8,0,0)]:
    dz = torch.zeros(M, N, device="cuda"); x = torch.zeros(M, K, device="cuda")
    dz[m0, n0] = 1.0; x[m0, k0] = 2.0
    W, S = run(dz, x, N, K)
    nz = torch.nonzero(W != 0)
    print((m0, n0, k0), "slabs", S, "nonzeros:", [(int(i), int(j), float(W[i, j])) for i, j in nz[:8]], "nan:", int(torch.isnan(W).sum()))
