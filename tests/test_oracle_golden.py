"""CPU: pin the oracle against fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  Integer/byte results bit-exact; floating point
within rel 1e-5 (the tolerance BASELINE.json's north_star states)."""
import numpy as np
import pytest

from oracle import ddpg as OD
from oracle import her as OH
from tests.helpers import (DDPG_LARGE_CASES, TD3_LARGE_CASES, golden_update_inputs, weight_error_report,
                           DDPG_CASES, HER_CASES, bits, ddpg_params_from_golden, her_episodes,
                           load, weights_close)

RTOL = 1e-5


def test_reward_known_answers_bit_exact():
    g = load("reward_kat")
    r = OH.compute_reward(g["a"], g["b"])
    assert np.array_equal(bits(r), bits(g["r"]))
    # success is -0.0 (sign bit set), failure -1.0
    assert set(np.unique(bits(r))) == {0x80000000, 0xBF800000}


def _replay_into_oracle(g):
    O, G, A, k, max_len, nenvs = (int(x) for x in g["meta"])
    eps = her_episodes(g)
    draws = []
    for ep in eps:
        T = ep["s"].shape[0]
        draws.extend(int(v) for v in ep["fut"][:T - 1].reshape(-1))
    it = iter(draws)
    buf = OH.HERBufferOracle(max_len, 50, 1, k_future=k, randint=lambda a, b: next(it))
    for ep in eps:
        T = ep["s"].shape[0]
        for t in range(T):
            buf.push(0, ep["s"][t], ep["a"][t], ep["ns"][t], ep["r"][t], bool(ep["d"][t]),
                     ep["dg"][t], ep["ag"][t])
    assert next(it, None) is None
    return buf, eps, k


@pytest.mark.parametrize("case", HER_CASES)
def test_her_oracle_matches_reference_dump(case):
    g = load("her_" + case)
    buf, eps, k = _replay_into_oracle(g)
    assert len(buf) == int(g["len"])
    s, a, r, ns, d = OH.collate(list(buf.buffer))
    for got, key in ((s, "dump_s"), (a, "dump_a"), (ns, "dump_ns")):
        assert np.array_equal(bits(got), bits(g[key])), key
    assert np.array_equal(bits(r[:, 0]), bits(g["dump_r"]))
    assert np.array_equal(bits(d[:, 0]), bits(g["dump_d"]))


@pytest.mark.parametrize("case", HER_CASES)
def test_her_oracle_sample_matches_reference(case):
    g = load("her_" + case)
    buf, _, _ = _replay_into_oracle(g)
    for bi in range(int(g["n_batches"])):
        idx = g[f"b{bi}_idx"]
        buf._sample = lambda population, B, _i=idx: [population[int(i)] for i in _i]
        out = buf.sample(len(idx))
        for got, key in zip(out, ("s", "a", "r", "ns", "d")):
            assert np.array_equal(bits(got), bits(g[f"b{bi}_{key}"])), (bi, key)


@pytest.mark.parametrize("case", HER_CASES)
def test_vectorised_materialise_equals_eager(case):
    g = load("her_" + case)
    eps = her_episodes(g)
    k = int(g["meta"][3])
    parts = [OH.materialise_episode(e["s"], e["a"], e["ns"], e["r"], e["d"], e["ag"], e["fut"], k)
             for e in eps]
    full = [np.concatenate([p[i] for p in parts]) for i in range(5)]
    n = int(g["len"])
    for got, key in zip(full, ("dump_s", "dump_a", "dump_r", "dump_ns", "dump_d")):
        ref = g[key]
        got = got[-n:].reshape(ref.shape)
        assert np.array_equal(bits(got), bits(ref)), key


def test_sample_underfilled_asserts():
    buf = OH.HERBufferOracle(100, 50, 1)
    with pytest.raises(AssertionError):
        buf.sample(1)


def test_normalizer_oracle_matches_reference():
    g = load("normalizer")
    for tag, dim in (("obs", 19), ("dg", 3)):
        nz = OH.RunningNormalizerOracle(dim)
        for i in range(6):
            nz.update(g[f"{tag}_x{i}"])
            np.testing.assert_allclose(nz.mean, g[f"{tag}_mean{i}"], rtol=1e-12, atol=1e-14)
            np.testing.assert_allclose(nz.var, g[f"{tag}_var{i}"], rtol=1e-12, atol=1e-14)
            assert nz.count == pytest.approx(float(g[f"{tag}_count{i}"]), rel=1e-15)
        np.testing.assert_allclose(nz.normalize(g[f"{tag}_q"]), g[f"{tag}_qn"], rtol=1e-12)


def make_ddpg_oracle(g):
    D, A, H, L, B, seed, ac_T, cr_T, freq = (int(x) for x in g["meta"])
    gamma, tau, clip, alr, clr, alr_min, clr_min = (float(x) for x in g["hp"])
    rng = np.random.default_rng(seed)
    actor = OD.init_mlp(rng, D, H, A, L)
    critic = OD.init_mlp(rng, D + A, H, 1, L)
    orc = OD.DDPGOracle(actor, critic, gamma=gamma, tau=tau, grad_clip=clip, actor_lr=alr,
                        critic_lr=clr, actor_lr_min=alr_min, critic_lr_min=clr_min,
                        ac_scheduler_steps=ac_T, cr_scheduler_steps=cr_T, ac_update_freq=freq)
    return orc, rng


def ddpg_batch(g, si):
    return tuple(g[f"s{si}_batch_{k}"] for k in ("s", "a", "r", "ns", "d"))


@pytest.mark.parametrize("case", DDPG_CASES)
def test_ddpg_oracle_matches_reference(case):
    g = load("ddpg_" + case)
    orc, _ = make_ddpg_oracle(g)
    steps = g["steps"]
    for si, step in enumerate(steps):
        info = orc.update_on_batch(int(step), *ddpg_batch(g, si))
        ref = g[f"s{si}_info"]
        assert len(info) == len(ref)
        np.testing.assert_allclose(np.array(info), ref, rtol=2e-5, atol=1e-6)
        np.testing.assert_allclose([orc.critic_sched.lr, orc.actor_sched.lr], g[f"s{si}_lr"],
                                   rtol=1e-12)
        if f"s{si}_actor.base_net.0.weight" in g.files:
            for tag, params in (("actor", orc.actor), ("critic", orc.critic),
                                ("target_actor", orc.target_actor),
                                ("target_critic", orc.target_critic)):
                ref_p = ddpg_params_from_golden(g, si, tag)
                lr = max(float(g["hp"][3]), float(g["hp"][4]))
                for (w, b), (rw, rb) in zip(params, ref_p):
                    assert weights_close(w, rw, lr, si + 1), (si, tag)
                    assert weights_close(b, rb, lr, si + 1), (si, tag)


@pytest.mark.parametrize("case", DDPG_LARGE_CASES)
def test_ddpg_oracle_matches_reference_large_batch(case):
    """Batches of 2048 / 8192 on the PickAndPlace shape (regenerated from the seed): the oracle against the
    unmodified reference's metrics (every step) and weights (last step).  fp32 batch means over B terms: BLAS
    and NumPy sum in different orders, the metric tolerance grows like sqrt(B / 256)."""
    g = load("ddpg_" + case)
    orc, _ = make_ddpg_oracle(g)
    _, batches = golden_update_inputs(g)
    B = int(g["meta"][4])
    rtol = 2e-5 * max(1.0, (B / 256.0) ** 0.5)
    n = len(g["steps"])
    for si, step in enumerate(g["steps"]):
        info = orc.update_on_batch(int(step), *batches[si])
        np.testing.assert_allclose(np.array(info), g[f"s{si}_info"], rtol=rtol, atol=1e-6)
    lr = max(float(g["hp"][3]), float(g["hp"][4]))
    for tag, params in (("actor", orc.actor), ("critic", orc.critic), ("target_actor", orc.target_actor),
                        ("target_critic", orc.target_critic)):
        for (w, b), (rw, rb) in zip(params, ddpg_params_from_golden(g, n - 1, tag)):
            assert weights_close(w, rw, lr, n) and weights_close(b, rb, lr, n), (tag, weight_error_report(w, rw, lr, n))


def test_checkpoint_forward_matches_reference():
    g = load("checkpoint_reach")
    actor = OD.state_dict_to_params({k[6:]: g[k] for k in g.files if k.startswith("actor.")},
                                    "base_net")
    critic = OD.state_dict_to_params({k[7:]: g[k] for k in g.files if k.startswith("critic.")},
                                     "net")
    act, _ = OD.mlp_forward(actor, g["x"], final_tanh=True)
    q, _ = OD.mlp_forward(critic, np.concatenate([g["x"], act], -1), final_tanh=False)
    np.testing.assert_allclose(act, g["act"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(q, g["q"], rtol=1e-5, atol=1e-6)


# ---- TD3 ------------------------------------------------------------------------------------
TD3_CASES = ["push_h64", "pickplace_h256"]


def make_td3_oracle(g):
    from oracle import td3 as OT
    D, A, H, L, B, seed, freq = (int(x) for x in g["meta"])
    gamma, tau, clip, lr, pn, nc = (float(x) for x in g["hp"])
    rng = np.random.default_rng(seed)
    nets = [OD.init_mlp(rng, D, H, A, L), OD.init_mlp(rng, D + A, H, 1, L), OD.init_mlp(rng, D + A, H, 1, L)]
    return OT.TD3Oracle(*nets, gamma=gamma, tau=tau, grad_clip=clip, actor_lr=lr, critic_lr=lr,
                        policy_noise=pn, noise_clamp=nc, ac_update_freq=freq), nets


@pytest.mark.parametrize("case", TD3_CASES + TD3_LARGE_CASES)
def test_td3_oracle_matches_reference(case):
    g = load("td3_" + case)
    orc, _ = make_td3_oracle(g)
    lr = float(g["hp"][3])
    n = len(g["steps"])
    _, batches = golden_update_inputs(g, n_critics=2)
    rtol = 2e-5 * max(1.0, (int(g["meta"][4]) / 256.0) ** 0.5)
    for si, step in enumerate(g["steps"]):
        info = orc.update_on_batch(int(step), *batches[si], g[f"s{si}_noise"])
        ref = g[f"s{si}_info"]
        assert len(info) == len(ref)
        np.testing.assert_allclose(np.array(info), ref, rtol=rtol, atol=1e-6)
    for tag in ("actor", "critic_1", "critic_2", "target_actor", "target_critic_1", "target_critic_2"):
        for (w, b), (rw, rb) in zip(getattr(orc, tag), ddpg_params_from_golden(g, n - 1, tag)):
            assert weights_close(w, rw, lr, n) and weights_close(b, rb, lr, n), tag
