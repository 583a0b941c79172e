"""GPU parity: the SAC / TQC update (csrc/sac.cu through gcrl_sac_*) versus fixtures dumped from the
unmodified reference ``SACAgent.update`` / ``TQCAgent.update`` (src/agent.py:659-699, :1062-1100)
with the recorded ``Normal.rsample`` noise, and versus the NumPy oracle on ragged shapes.

Tolerance: metrics rel 2e-5 + abs 1e-6; weights tests.helpers.weights_close /
assert_sac_actor_close (pre-BatchNorm Linear biases only to Adam's hard bound -- their true
gradient is zero, see the helper); BatchNorm running variance rel 1e-5."""
import os

import numpy as np
import pytest

from tests.helpers import (SAC_CASES, assert_sac_actor_close, load, sac_initial_nets, sac_oracle_from_golden,
                           sac_params_from_golden, weights_close)
from tests.test_ddpg_gpu import batch_to_device, make_config

pytestmark = pytest.mark.gpu


def sac_config(**over):
    cfg = make_config(**over)
    cfg.alpha_lr = over.get("alpha_lr", 3e-4)
    cfg.alpha_min = 0.05
    cfg.alpha_min_steps = over.get("alpha_min_steps", 2)
    return cfg


def load_initial(ag, actor0, stats0, critics0):
    L = ag.config.layer_count
    for l in range(L):
        ag.actor.set_linear(l, *actor0[2 * l])
        ag.actor.set_bn(l, actor0[2 * l + 1][0], actor0[2 * l + 1][1], stats0[l][0], stats0[l][1])
    ag.actor.set_linear(L, *actor0[2 * L])
    ag.actor.set_linear(L + 1, *actor0[2 * L + 1])
    for view, p in zip(ag._critic_views, critics0):
        view.set_layers(p)
    ag.update_target_network()


def actor_params(ag):
    L = ag.config.layer_count
    params, stats = [], []
    for l in range(L):
        params.append(list(ag.actor.linear(l)))
        g, be, rm, rv = ag.actor.bn(l)
        params.append([g, be])
        stats.append([rm, rv])
    params.append(list(ag.actor.linear(L)))
    params.append(list(ag.actor.linear(L + 1)))
    return params, stats


def make_agent(algo, g, max_batch=None):
    from gcrl_b200 import SACAgent, TQCAgent
    D, A, H, L, B, seed, freq, gstep, amin = (int(x) for x in g["meta"])
    gamma, tau, clip, lr, alpha_lr = (float(x) for x in g["hp"])
    cfg = sac_config(hidden_dim=H, layer_count=L, batch_size=B, gamma=gamma, tau=tau, grad_clip=clip, actor_lr=lr,
                     critic_lr=lr, actor_lr_min=lr, critic_lr_min=lr, ac_update_freq=freq, alpha_lr=alpha_lr,
                     alpha_min_steps=amin)
    ag = (SACAgent if algo == "sac" else TQCAgent)(D, A, cfg, None, 1, gstep)
    load_initial(ag, *sac_initial_nets(algo, g))
    return ag


@pytest.mark.parametrize("algo,case", SAC_CASES)
def test_update_matches_reference_fixture(algo, case):
    import torch
    g = load(f"{algo}_{case}")
    ag = make_agent(algo, g)
    lr = float(g["hp"][3])
    steps = [int(x) for x in g["steps"]]
    for si, step in enumerate(steps):
        info = ag.update(step, batch=batch_to_device(g, si), eps_next=torch.from_numpy(g[f"s{si}_eps_next"]).cuda(),
                         eps_cur=torch.from_numpy(g[f"s{si}_eps_cur"]).cuda())
        ref = g[f"s{si}_info"]
        assert len(info) == len(ref), "tuple arity (9 with the actor step, else 6)"
        np.testing.assert_allclose(np.array([float(x) for x in info]), ref, rtol=2e-5, atol=1e-6)
        np.testing.assert_allclose(ag.get_log_alpha(), float(g[f"s{si}_log_alpha"][0]), rtol=1e-5, atol=1e-8)
    si, n = len(steps) - 1, len(steps)
    params, stats = actor_params(ag)
    ref = sac_params_from_golden(g, si, "actor")
    assert_sac_actor_close(params, ref["params"], lr, n)
    for (m, v), (rm, rv) in zip(stats, ref["stats"]):
        np.testing.assert_allclose(m, rm, rtol=1e-5, atol=2.0 * lr * n)    # carries the noise-driven bias
        np.testing.assert_allclose(v, rv, rtol=1e-5, atol=1e-6)
    for i in (0, ag.N_CRITICS - 1):
        tag = f"critic_{i + 1}" if algo == "sac" else f"critic_{i}"
        for view, t in ((ag._critic_views[i], tag), (ag._target_views[i], "target_" + tag)):
            for (w, b), (rw, rb) in zip(view.layers(), sac_params_from_golden(g, si, t)):
                assert weights_close(w, rw, lr, n) and weights_close(b, rb, lr, n), (case, t)
    np.testing.assert_allclose(ag.select_action(g["eval_x"], eval_action=True), g["eval_act"], rtol=1e-5,
                               atol=0.1 * lr * n)


@pytest.mark.parametrize("algo,B,H,L,D,A", [("sac", 33, 100, 2, 22, 3), ("tqc", 777, 64, 1, 7, 1),
                                            ("tqc", 1000, 256, 3, 23, 4), ("sac", 2, 64, 3, 10, 3)])
def test_update_matches_oracle_odd_shapes(algo, B, H, L, D, A):
    """Ragged shapes (B not a tile multiple, B = 2, H not a multiple of 32, A in 1..4)."""
    import torch
    from gcrl_b200 import SACAgent, TQCAgent
    from oracle import ddpg as OD
    from oracle import sac as OS
    rng = np.random.default_rng(B * 7 + H)
    cfg = sac_config(hidden_dim=H, layer_count=L, batch_size=B, grad_clip=0.5, tau=0.05, alpha_min_steps=1,
                     alpha_lr=1e-2)
    n = 2 if algo == "sac" else 5
    actor0, stats0 = OS.init_sac_actor(rng, D, H, A, L, head_scale=0.1, log_std_bias=-1.0)
    critics0 = [OD.init_mlp(rng, D + A, H, 1, L) for _ in range(n)]
    ag = (SACAgent if algo == "sac" else TQCAgent)(D, A, cfg, None, 1, 2)
    load_initial(ag, actor0, stats0, critics0)
    orc = OS.SACOracle(algo, actor0, stats0, critics0, act_dim=A, gamma=cfg.gamma, tau=cfg.tau,
                       grad_clip=cfg.grad_clip, actor_lr=cfg.actor_lr, critic_lr=cfg.critic_lr, alpha_lr=1e-2,
                       alpha_min_steps=1, gradient_step=2)
    for step in (1, 2, 3):
        s = rng.standard_normal((B, D)).astype(np.float32)
        ns = (s + 0.1 * rng.standard_normal((B, D))).astype(np.float32)
        a = rng.uniform(-1, 1, (B, A)).astype(np.float32)
        r = -(rng.random((B, 1)) > 0.3).astype(np.float32)
        d = (rng.random((B, 1)) < 0.1).astype(np.float32)
        e1 = rng.standard_normal((B, A)).astype(np.float32)
        e2 = rng.standard_normal((B, A)).astype(np.float32)
        want = np.array(orc.update_on_batch(step, s, a, r, ns, d, e1, e2), np.float64)
        got = ag.update(step, batch=tuple(torch.from_numpy(x).cuda() for x in (s, a, r, ns, d)),
                        eps_next=torch.from_numpy(e1).cuda(), eps_cur=torch.from_numpy(e2).cuda())
        rtol = 5e-5 * max(1.0, (B / 256.0) ** 0.5)
        np.testing.assert_allclose(np.array([float(x) for x in got]), want, rtol=rtol, atol=2e-6)
        np.testing.assert_allclose(ag.get_log_alpha(), float(orc.log_alpha), rtol=1e-5, atol=1e-8)


def test_sac_from_buffer_own_noise_and_checkpoint_files(tmp_path):
    from gcrl_b200 import SACAgent, TQCAgent
    from tests.helpers import her_episodes
    g = load("her_reach_small")
    for cls, files in ((SACAgent, ["actor.pth", "critic_1.pth", "critic_2.pth", "log_alpha.pth"]),
                       (TQCAgent, ["actor.pth"] + [f"critic_{i}.pth" for i in range(5)] + ["log_alpha.pth"])):
        ag = cls(10, 3, sac_config(batch_size=64, max_len=100000), None, 2, 2)
        for ep in her_episodes(g):
            ag.buffer.push_episode(ep["s"], ep["a"], ep["ns"], ep["r"], ep["d"], ep["ag"], ep["fut"])
        for step in (1, 2, 3):
            info = ag.update(step)
            assert len(info) == 9 and all(np.isfinite(float(x)) for x in info)
        assert np.isfinite(ag.alpha.item())
        act = ag.select_action(np.zeros((4, 10), np.float32))
        assert act.shape == (4, 3) and np.all(np.abs(act) <= 1)
        out = tmp_path / cls.__name__
        ag.save_weights(str(out))
        assert sorted(os.listdir(out)) == sorted(files)
        ag2 = cls(10, 3, sac_config(batch_size=64, max_len=100000), str(out), 2, 2)
        x = np.random.default_rng(0).standard_normal((5, 10)).astype(np.float32)
        assert np.array_equal(ag.select_action(x, eval_action=True), ag2.select_action(x, eval_action=True))
        import torch
        sd = torch.load(str(out / "actor.pth"))
        assert "base_net.1.running_mean" in sd and "mean_head.weight" in sd and "log_std_head.bias" in sd
