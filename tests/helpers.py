"""Helpers shared by the parity tests (golden loading, bit views)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

HER_CASES = ["reach_small", "push_evict", "pickplace_k8", "k0"]
DDPG_CASES = ["reach_h64", "push_h256", "pickplace_l2_cosine"]
# large-batch updates on the PickAndPlace shape (the tensor-core engine's batch sizes); the batches are
# regenerated from the seed (golden_update_inputs), only the reference's outputs are stored
DDPG_LARGE_CASES = ["pickplace_B2048", "pickplace_B8192"]
TD3_LARGE_CASES = ["pickplace_B4096"]
SAC_CASES = [("sac", "push_h64"), ("sac", "pickplace_h256"), ("tqc", "slide_h64"), ("tqc", "push_h256")]


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


def golden_update_inputs(g, n_critics=1):
    """-> (initial networks [actor, critic, ...] and the per-step batches) of a DDPG / TD3 update fixture, drawn
    exactly as tests/golden/make_golden.py::ddpg_case / td3_case drew them: seeded generator, init_mlp for the
    actor and every critic, then per step s, ns, a, r, d.  Stored batches are returned as stored; fixtures
    written with store_batches=False carry a float64 checksum per tensor that pins the regenerated stream."""
    from oracle import ddpg as OD
    D, A, H, L, B, seed = (int(x) for x in g["meta"][:6])
    rng = np.random.default_rng(seed)
    nets = [OD.init_mlp(rng, D, H, A, L)] + [OD.init_mlp(rng, D + A, H, 1, L) for _ in range(n_critics)]
    batches = []
    for si in range(len(g["steps"])):
        s = rng.standard_normal((B, D)).astype(np.float32)
        ns = (s + 0.1 * rng.standard_normal((B, D))).astype(np.float32)
        a = rng.uniform(-1, 1, (B, A)).astype(np.float32)
        r = -(rng.random((B, 1)) > 0.3).astype(np.float32)
        d = (rng.random((B, 1)) < 0.1).astype(np.float32)
        if f"s{si}_batch_s" in g.files:
            stored = tuple(g[f"s{si}_batch_{k}"] for k in ("s", "a", "r", "ns", "d"))
            for x, y in zip((s, a, r, ns, d), stored):
                assert np.array_equal(x, y), "seeded stream differs from the stored batch"
        else:
            got = np.array([x.astype(np.float64).sum() for x in (s, a, r, ns, d)])
            assert np.array_equal(got, g[f"s{si}_batch_sum"]), "regenerated batch does not match the fixture's checksum"
        batches.append((s, a, r, ns, d))
    return nets, batches


def weight_error_report(w, ref, lr, nsteps, rtol=1e-5):
    """(max abs error, 99.99-percentile, the bulk allowance of weights_close, fraction of it used) -- printed by
    the parity tests so that a reader sees how much of the stated tolerance is actually consumed."""
    err = np.abs(np.asarray(w, np.float64) - np.asarray(ref, np.float64)).ravel()
    tol = rtol * float(np.max(np.abs(ref))) + 5e-3 * lr * nsteps
    p9999 = float(np.quantile(err, 0.9999)) if err.size else 0.0
    return float(err.max()) if err.size else 0.0, p9999, tol, (p9999 / tol if tol else 0.0)


def rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-30))


def weights_close(w, ref, lr, nsteps, rtol=1e-5, outlier_frac=2e-4, extra=None):
    """Stated fp32 tolerance for post-update weights.

    Bulk: |w - ref| <= rtol * max|ref| + 5e-3 * lr * nsteps element-wise, for all but a fraction
    ``outlier_frac`` of the elements; the outliers stay inside Adam's hard bound 2 * lr * nsteps.
    The first term is the north-star's rel 1e-5 (norm-wise per tensor).  The rest covers Adam's
    sign regime: the step is lr * m / (sqrt(v) + 1e-8), i.e. ~ lr * sign(g) in the first steps,
    so for the few gradient entries whose magnitude is at the fp32 summation-order noise of the
    batch reduction (|g| <~ 1e-6 relative to the tensor) the quotient m / sqrt(v) is decided by
    that noise and can move by up to its full range.  Observed against the unmodified reference:
    1 element in 65 536 at 3.3e-3 * lr, all others < 1e-7 absolute; the NumPy oracle shows the
    same effect against torch.  Gradients, losses and Q values carry no such amplification and
    are held to rel 2e-5.

    ``extra``: optional per-ELEMENT additional allowance (same shape as the tensor), e.g. the oracle's bound for
    the elements a LeakyReLU sign flip of a near-zero pre-activation can move
    (oracle/ddpg.py::DDPGOracle._flip_track); elements it leaves at 0 stay under the tolerance above.
    """
    w = np.asarray(w, np.float64)
    ref = np.asarray(ref, np.float64)
    err = np.abs(w - ref)
    tol = rtol * np.max(np.abs(ref)) + 5e-3 * lr * nsteps
    if extra is not None:
        tol = tol + 1.5 * np.asarray(extra, np.float64).reshape(ref.shape)
    if float(np.max(err)) > 2.0 * lr * nsteps + rtol * np.max(np.abs(ref)):
        return False
    return bool(np.count_nonzero(err > tol) <= outlier_frac * err.size)


def sac_params_from_golden(g, si, tag):
    """SACActorModel state_dict (src/model.py:100-116: base_net.{3l} Linear, base_net.{3l+1}
    BatchNorm1d, mean_head, log_std_head) -> oracle layout; Critic -> list of [W, b]."""
    pref = f"s{si}_{tag}."
    if tag != "actor":
        return ddpg_params_from_golden(g, si, tag)
    L = len([k for k in g.files if k.startswith(pref + "base_net.") and k.endswith("running_mean")])
    params, stats = [], []
    for l in range(L):
        params.append([g[f"{pref}base_net.{3 * l}.weight"], g[f"{pref}base_net.{3 * l}.bias"]])
        params.append([g[f"{pref}base_net.{3 * l + 1}.weight"], g[f"{pref}base_net.{3 * l + 1}.bias"]])
        stats.append([g[f"{pref}base_net.{3 * l + 1}.running_mean"], g[f"{pref}base_net.{3 * l + 1}.running_var"]])
    params.append([g[pref + "mean_head.weight"], g[pref + "mean_head.bias"]])
    params.append([g[pref + "log_std_head.weight"], g[pref + "log_std_head.bias"]])
    return {"params": params, "stats": stats}


def sac_initial_nets(algo, g):
    """Seeded initial parameters exactly as tests/golden/make_golden.py::sac_case drew them."""
    from oracle import ddpg as OD
    from oracle import sac as OS
    D, A, H, L, B, seed = (int(x) for x in g["meta"][:6])
    rng = np.random.default_rng(seed)
    actor0, stats0 = OS.init_sac_actor(rng, D, H, A, L, head_scale=0.1, log_std_bias=-1.0)
    critics0 = [OD.init_mlp(rng, D + A, H, 1, L) for _ in range(2 if algo == "sac" else 5)]
    return actor0, stats0, critics0


def sac_oracle_from_golden(algo, g):
    from oracle import sac as OS
    D, A, H, L, B, seed, freq, gstep, amin = (int(x) for x in g["meta"])
    gamma, tau, clip, lr, alpha_lr = (float(x) for x in g["hp"])
    actor0, stats0, critics0 = sac_initial_nets(algo, g)
    return OS.SACOracle(algo, actor0, stats0, critics0, act_dim=A, gamma=gamma, tau=tau, grad_clip=clip,
                        actor_lr=lr, critic_lr=lr, alpha_lr=alpha_lr, alpha_min_steps=amin,
                        gradient_step=gstep, ac_update_freq=freq)


def assert_sac_actor_close(params, ref_params, lr, nsteps):
    """SAC actor parameters in oracle layout.  The bias of a Linear that feeds BatchNorm has an
    exactly-zero true gradient (the batch mean is subtracted again), so what reaches AdamW is pure
    fp32 rounding noise, which Adam normalises to +-lr steps in a random direction: those biases
    (which cannot change the network's output) are only held to Adam's hard bound."""
    L = (len(params) - 2) // 2
    for i, ((w, b), (rw, rb)) in enumerate(zip(params, ref_params)):
        assert weights_close(w, rw, lr, nsteps), ("actor", i, rel_err(w, rw))
        if i < 2 * L and i % 2 == 0:
            assert float(np.max(np.abs(np.asarray(b, np.float64) - rb))) <= 2.0 * lr * nsteps + 1e-6, ("actor bias", i)
        else:
            assert weights_close(b, rb, lr, nsteps), ("actor", i, rel_err(b, rb))


# ---- prioritised replay fixtures (tests/golden/make_golden.py::per_case) --------------------------------
PER_CASES = [("ddpg", "reach_h64"), ("ddpg", "push_h256"), ("td3", "push_h64"), ("sac", "push_h64"),
             ("tqc", "slide_h64")]


def per_meta(g):
    keys = ("D", "A", "H", "L", "B", "seed", "freq", "gstep", "max_len", "n0", "push_per_step", "beta_end")
    m = {k: int(v) for k, v in zip(keys, g["meta"])}
    m.update({k: float(v) for k, v in zip(("gamma", "tau", "clip", "lr", "alpha", "beta"), g["hp"])})
    return m


def per_pushes(g, si):
    """Row range [lo, hi) of the recorded push stream that the generator appended before step si."""
    m = per_meta(g)
    lo = 0 if si == 0 else m["n0"] + (si - 1) * m["push_per_step"]
    return lo, m["n0"] + si * m["push_per_step"]


def per_initial_nets(algo, g):
    """Seeded initial parameters exactly as per_case drew them (before the push stream)."""
    from oracle import ddpg as OD
    from oracle import sac as OS
    m = per_meta(g)
    rng = np.random.default_rng(m["seed"])
    D, A, H, L = m["D"], m["A"], m["H"], m["L"]
    if algo in ("ddpg", "td3"):
        nets = [OD.init_mlp(rng, D, H, A, L)]
        nets += [OD.init_mlp(rng, D + A, H, 1, L) for _ in range(1 if algo == "ddpg" else 2)]
        return nets
    actor0, stats0 = OS.init_sac_actor(rng, D, H, A, L, head_scale=0.1, log_std_bias=-1.0)
    critics0 = [OD.init_mlp(rng, D + A, H, 1, L) for _ in range(2 if algo == "sac" else 5)]
    return actor0, stats0, critics0


def per_oracle(algo, g):
    """The update oracle of a PER fixture; ``step(step, batch, weights, normals)`` -> info tuple."""
    from oracle import ddpg as OD
    from oracle import sac as OS
    from oracle import td3 as OT
    m = per_meta(g)
    kw = dict(gamma=m["gamma"], tau=m["tau"], grad_clip=m["clip"], actor_lr=m["lr"], critic_lr=m["lr"],
              ac_update_freq=m["freq"])
    nets = per_initial_nets(algo, g)
    if algo == "ddpg":
        orc = OD.DDPGOracle(*nets, **kw)
        return orc, lambda step, b, w, nrm: orc.update_on_batch(step, *b, weights=w)
    if algo == "td3":
        orc = OT.TD3Oracle(*nets, policy_noise=0.2, noise_clamp=0.5, **kw)
        return orc, lambda step, b, w, nrm: orc.update_on_batch(step, *b, nrm[0], weights=w)
    orc = OS.SACOracle(algo, *nets, act_dim=m["A"], alpha_lr=1e-2, alpha_min_steps=0, gradient_step=m["gstep"], **kw)
    return orc, lambda step, b, w, nrm: orc.update_on_batch(step, *b, nrm[0], nrm[1] if len(nrm) > 1 else None,
                                                            weights=w)


def per_td_position(algo, info_len):
    return {"ddpg": 2 if info_len == 6 else 1, "td3": 3 if info_len == 8 else 2}.get(algo, 3 if info_len == 9 else 2)


def ulp_diff_f32(a, b):
    """Distance in float32 units in the last place (same-sign finite values)."""
    a = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    return np.abs(a - b)
