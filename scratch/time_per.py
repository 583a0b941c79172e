"""Timing of PERBuffer.sample at N entries (CUDA events, L2 not flushed): benign vs extreme priority ratios."""
import sys, os, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "goal-conditioned-rl-framework_b200"))
from gcrl_b200 import PERBuffer
from gcrl_b200._lib import vp
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
rng = np.random.default_rng(0)
D, A = 21, 3
buf = PERBuffer(N, 0.6)
for lo in range(0, N, 1 << 16):
    n = min(1 << 16, N - lo)
    buf.push_rows(rng.standard_normal((n, D)).astype(np.float32), rng.uniform(-1, 1, (n, A)).astype(np.float32),
                  np.zeros(n, np.float32), rng.standard_normal((n, D)).astype(np.float32), np.zeros(n, np.float32))
out = buf._outputs(B)
w = torch.empty(B, device="cuda")
for name, prio in (("uniform 1.0", np.ones(N, np.float32)),
                   ("td-like (|td|+1e-6)^0.6, td ~ |N(0,0.3)|", ((np.abs(rng.normal(0, 0.3, N)) + 1e-6) ** 0.6).astype(np.float32)),
                   ("wide 1e-3..5", (rng.random(N) ** 3 * 5 + 1e-3).astype(np.float32))):
    buf.set_priorities(prio)
    u = rng.random(B)
    for _ in range(3):
        buf.sample_into(B, 0.4, out, vp(w.data_ptr()), u)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    it = 20
    e0.record()
    for _ in range(it):
        buf.sample_into(B, 0.4, out, vp(w.data_ptr()), u)
    e1.record(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(it):
        buf.sample_into(B, 0.4, out, vp(w.data_ptr()), u)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / it
    print(f"N={N} B={B} {name}: {e0.elapsed_time(e1) / it * 1e3:.1f} us/sample (device), {wall * 1e6:.1f} us wall, sequential={buf.last_sample_info()[1]}")
# the reference's own path on the host for scale
P = prio.copy(); t0 = time.perf_counter()
for _ in range(3):
    Pn = P / P.sum(); idx = np.random.choice(N, B, p=Pn)
print(f"numpy choice alone at N={N}: {(time.perf_counter() - t0) / 3 * 1e3:.2f} ms")
