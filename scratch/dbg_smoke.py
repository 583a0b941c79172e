import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "goal-conditioned-rl-framework_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from gcrl_b200 import HERBuffer
from oracle import her as OH
rng = np.random.default_rng(0)
O, G, A, k, T = 18, 3, 3, 4, 50
buf = HERBuffer(100000, 50, 1, k_future=k)
parts = []
for _ in range(12):
    obs = rng.standard_normal((T + 1, O)).astype(np.float32)
    ag = np.cumsum(rng.normal(0, 0.02, (T + 1, G)) * (rng.random((T + 1, 1)) > 0.3), 0).astype(np.float32)
    dg = rng.uniform(-0.15, 0.15, (1, G)).astype(np.float32)
    s = np.concatenate([obs[:-1], np.repeat(dg, T, 0)], -1)
    ns = np.concatenate([obs[1:], np.repeat(dg, T, 0)], -1)
    a = rng.uniform(-1, 1, (T, A)).astype(np.float32)
    r = OH.compute_reward(ag[1:], np.repeat(dg, T, 0))
    d = np.zeros(T, np.float32)
    fut = np.zeros((T, k), np.uint8)
    for t in range(T - 1):
        fut[t] = rng.integers(t + 1, T, k)
    buf.push_episode(s, a, ns, r, d, ag[1:], fut)
    parts.append(OH.materialise_episode(s, a, ns, r, d, ag[1:], fut, k))
full = [np.concatenate([p[i] for p in parts]) for i in range(5)]
n = len(buf)
idx = np.arange(n)
got = buf.sample_host(n, indices=idx)
R, Rr = got[2][:, 0], full[2][:, 0]
bad = np.nonzero(R.view(np.uint32) != Rr.view(np.uint32))[0]
print("n", n, "bad rewards", len(bad))
for b in bad[:10]:
    o = b % 246
    print(b, "ep", b // 246, "o", o, "t", min(o // 5, 49), "j", o % 5 if o < 245 else 0, "gpu", R[b], "ref", Rr[b],
          np.signbit(R[b]), np.signbit(Rr[b]))
for i, nm in enumerate(("s", "a", "r", "ns", "d")):
    print(nm, np.array_equal(got[i].view(np.uint32), full[i].view(np.uint32).reshape(got[i].shape)))
