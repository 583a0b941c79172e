"""The SAC / TQC oracle (oracle/sac.py) against fixtures dumped from the unmodified reference
``SACAgent.update`` / ``TQCAgent.update`` (src/agent.py:659-699, :1062-1100) with recorded
``rsample`` noise.  Tolerance: metrics rel 2e-5 + abs 1e-6, weights tests.helpers.weights_close."""
import numpy as np
import pytest

from tests.helpers import SAC_CASES, load, sac_oracle_from_golden, sac_params_from_golden, weights_close, assert_sac_actor_close


@pytest.mark.parametrize("algo,case", SAC_CASES)
def test_oracle_matches_reference_fixture(algo, case):
    g = load(f"{algo}_{case}")
    orc = sac_oracle_from_golden(algo, g)
    lr = float(g["hp"][3])
    steps = [int(x) for x in g["steps"]]
    for si, step in enumerate(steps):
        b = [g[f"s{si}_batch_{k}"] for k in ("s", "a", "r", "ns", "d")]
        info = orc.update_on_batch(step, *b, g[f"s{si}_eps_next"], g[f"s{si}_eps_cur"])
        ref = g[f"s{si}_info"]
        assert len(info) == len(ref)
        np.testing.assert_allclose(np.array(info, np.float64), ref, rtol=2e-5, atol=1e-6)
        np.testing.assert_allclose(float(orc.log_alpha), float(g[f"s{si}_log_alpha"][0]), rtol=1e-5, atol=1e-8)
    si = len(steps) - 1
    got = {"actor": (orc.actor, orc.actor_stats)}
    n = orc.n
    for i in (0, n - 1):
        tag = f"critic_{i + 1}" if algo == "sac" else f"critic_{i}"
        got[tag] = orc.critics[i]
        got["target_" + tag] = orc.target_critics[i]
    for tag, val in got.items():
        ref = sac_params_from_golden(g, si, tag)
        if tag == "actor":
            params, stats = val
            assert_sac_actor_close(params, ref["params"], lr, len(steps))
            for (m, v), (rm, rv) in zip(stats, ref["stats"]):
                # the running mean carries the (noise-driven, see assert_sac_actor_close) pre-BN bias
                np.testing.assert_allclose(m, rm, rtol=1e-5, atol=2.0 * lr * len(steps))
                np.testing.assert_allclose(v, rv, rtol=1e-5, atol=1e-6)
        else:
            for (w, b), (rw, rb) in zip(val, ref):
                assert weights_close(w, rw, lr, len(steps)) and weights_close(b, rb, lr, len(steps)), tag
    from oracle import sac as OS
    act, _, _ = OS.actor_sample(orc.actor, orc.actor_stats, g["eval_x"], None, train=False, deterministic=True)
    # eval mode normalises with running_mean - which lags the noise-driven pre-BN bias (above)
    np.testing.assert_allclose(act, g["eval_act"], rtol=1e-5, atol=0.1 * lr * len(steps))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sync_bn_exchange_equals_batchnorm_on_the_concatenated_batch(world):
    """The data-parallel BatchNorm exchange (csrc/sac.cu, restated in oracle/sac.py::sync_bn_*): per-rank (mean, M2)
    merged in rank order, and the input gradient from column sums over all ranks, against the single-batch formulas of
    the oracle's SACActorModel (src/model.py:103-111) on the concatenated rows.  The per-rank backward carries
    gradients `world` times larger than the concatenated run (local loss = mean over B, not world * B rows): the
    test removes that factor, as the gradient average over the ranks does."""
    from oracle import sac as OS
    rng = np.random.default_rng(world)
    B, H = 96, 40
    z = [(rng.standard_normal((B, H)) * 1.7 + rng.standard_normal(H) * 3).astype(np.float32) for _ in range(world)]
    zc = np.concatenate(z)
    mu_c = zc.mean(axis=0, dtype=np.float32)
    var_c = ((zc - mu_c) ** 2).mean(axis=0, dtype=np.float32)
    means = [x.mean(axis=0, dtype=np.float32) for x in z]
    m2s = [((x - m) ** 2).sum(axis=0, dtype=np.float32) for x, m in zip(z, means)]
    mu, var = OS.sync_bn_merge(means, m2s, B)
    np.testing.assert_allclose(mu, mu_c, rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(var, var_c, rtol=5e-6)
    # backward through y = relu(xhat * gamma + beta)
    gamma = rng.uniform(0.5, 1.5, H).astype(np.float32)
    invstd = (1.0 / np.sqrt(var + 1e-5)).astype(np.float32)
    xhat = [((x - mu) * invstd).astype(np.float32) for x in z]
    dy = [(rng.standard_normal((B, H)) * (rng.random((B, H)) > 0.4)).astype(np.float32) for _ in range(world)]
    s1 = [d.sum(0, dtype=np.float32) for d in dy]
    s2 = [(d * xh).sum(0, dtype=np.float32) for d, xh in zip(dy, xhat)]
    dz = [OS.sync_bn_backward(d, xh, invstd, gamma, s1, s2, world * B) for d, xh in zip(dy, xhat)]
    # one rank on the concatenated batch (actor_backward's BatchNorm step), upstream gradient / world
    dyc, xhc, n = np.concatenate(dy) / np.float32(world), np.concatenate(xhat), np.float32(world * B)
    dxh = (dyc * gamma).astype(np.float32)
    dz_c = invstd / n * (n * dxh - dxh.sum(0, dtype=np.float32) - xhc * (dxh * xhc).sum(0, dtype=np.float32))
    np.testing.assert_allclose(np.concatenate(dz) / world, dz_c, rtol=1e-4, atol=2e-7)
    # dgamma / dbeta: the average of the ranks' local sums is the concatenated batch's gradient
    np.testing.assert_allclose(sum(s2) / world, (dyc * xhc).sum(0), rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(sum(s1) / world, dyc.sum(0), rtol=1e-4, atol=1e-6)
