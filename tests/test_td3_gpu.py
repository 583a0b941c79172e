"""GPU parity: the TD3 update (same C entry points, algo = GCRL_ALGO_TD3) versus fixtures dumped
from the unmodified reference ``TD3Agent.update`` (src/agent.py:281-317) with its recorded
``torch.randn_like`` smoothing noise.  Tolerances as in test_ddpg_gpu.py."""
import numpy as np
import pytest

from oracle import ddpg as OD
from tests.helpers import (TD3_LARGE_CASES, ddpg_params_from_golden, golden_update_inputs, load, weight_error_report,
                           weights_close)
from tests.test_ddpg_gpu import batch_to_device, make_config

pytestmark = pytest.mark.gpu


def make_td3_from_golden(g):
    from gcrl_b200 import TD3Agent
    from gcrl_b200.agent import NET_ACTOR, NET_CRITIC, NET_CRITIC2
    D, A, H, L, B, seed, freq = (int(x) for x in g["meta"])
    gamma, tau, clip, lr, pn, nc = (float(x) for x in g["hp"])
    cfg = make_config(hidden_dim=H, layer_count=L, batch_size=B, gamma=gamma, tau=tau, grad_clip=clip,
                      actor_lr=lr, critic_lr=lr, actor_lr_min=lr, critic_lr_min=lr, policy_noise=pn,
                      noise_clamp=nc, ac_update_freq=freq)
    ag = TD3Agent(D, A, cfg, None, 1, 40)
    rng = np.random.default_rng(seed)
    ag._set_layers(NET_ACTOR, OD.init_mlp(rng, D, H, A, L))
    ag._set_layers(NET_CRITIC, OD.init_mlp(rng, D + A, H, 1, L))
    ag._set_layers(NET_CRITIC2, OD.init_mlp(rng, D + A, H, 1, L))
    ag.update_target_network()
    return ag


@pytest.mark.parametrize("case", ["push_h64", "pickplace_h256"])
def test_td3_update_matches_reference_fixture(case):
    import torch
    g = load("td3_" + case)
    ag = make_td3_from_golden(g)
    lr = float(g["hp"][3])
    n = len(g["steps"])
    for si, step in enumerate(g["steps"]):
        noise = torch.from_numpy(g[f"s{si}_noise"]).cuda()
        info = ag.update(int(step), batch=batch_to_device(g, si), noise=noise)
        ref = g[f"s{si}_info"]
        assert len(info) == len(ref), "tuple arity (8 with the actor step, else 6)"
        np.testing.assert_allclose(np.array([float(x) for x in info]), ref, rtol=2e-5, atol=1e-6)
    for tag, net in (("actor", ag.actor), ("critic_1", ag.critic_1), ("critic_2", ag.critic_2),
                     ("target_actor", ag.target_actor), ("target_critic_1", ag.target_critic_1),
                     ("target_critic_2", ag.target_critic_2)):
        for (w, b), (rw, rb) in zip(net.layers(), ddpg_params_from_golden(g, n - 1, tag)):
            assert weights_close(w, rw, lr, n) and weights_close(b, rb, lr, n), (case, tag)


@pytest.mark.parametrize("precision", [1, 0])
@pytest.mark.parametrize("case", TD3_LARGE_CASES)
def test_td3_large_batch_matches_reference_fixture(case, precision, capsys):
    """TD3Agent.update at batch 4096 on the PickAndPlace shape against the unmodified reference (recorded
    randn_like noise; batches regenerated from the seed): the default agent (hidden layers on tcgen05, 3xTF32)
    and the fp32 tile engine.  Metrics rel 2e-5 * sqrt(B / 256); weights of all six networks after the last step."""
    import torch
    from gcrl_b200 import TD3Agent
    from gcrl_b200.agent import NET_ACTOR, NET_CRITIC, NET_CRITIC2
    g = load("td3_" + case)
    D, A, H, L, B, seed, freq = (int(x) for x in g["meta"])
    gamma, tau, clip, lr, pn, nc = (float(x) for x in g["hp"])
    cfg = make_config(hidden_dim=H, layer_count=L, batch_size=B, gamma=gamma, tau=tau, grad_clip=clip,
                      actor_lr=lr, critic_lr=lr, actor_lr_min=lr, critic_lr_min=lr, policy_noise=pn,
                      noise_clamp=nc, ac_update_freq=freq)
    (actor0, c1, c2), batches = golden_update_inputs(g, n_critics=2)
    ag = TD3Agent(D, A, cfg, None, 1, 40, precision=precision)
    ag._set_layers(NET_ACTOR, actor0)
    ag._set_layers(NET_CRITIC, c1)
    ag._set_layers(NET_CRITIC2, c2)
    ag.update_target_network()
    # the tensor-core engine carries ~2e-6 relative error per 3xTF32 layer against ~5e-7 for the fp32 tiles
    # (profiles/README.md): its stated tolerance is 4x the fp32 engine's, for metrics and for weights alike
    eng = 4.0 if precision else 1.0
    rtol = eng * 2e-5 * max(1.0, (B / 256.0) ** 0.5)
    n = len(g["steps"])
    lines, failures = [f"td3 {case} precision={precision} (metric tolerance {rtol:.1e})"], []
    for si, step in enumerate(g["steps"]):
        noise = torch.from_numpy(g[f"s{si}_noise"]).cuda()
        info = ag.update(int(step), batch=tuple(torch.from_numpy(x).cuda() for x in batches[si]), noise=noise)
        got, ref = np.array([float(x) for x in info]), g[f"s{si}_info"]
        assert len(got) == len(ref), "tuple arity (8 with the actor step, else 6)"
        rel = np.abs(got - ref) / (np.abs(ref) + 1e-6)
        lines.append(f"  step {int(step)}: metric rel errors " + " ".join(f"{x:.1e}" for x in rel))
        if not np.allclose(got, ref, rtol=rtol, atol=1e-6):
            failures.append((int(step), got.tolist(), ref.tolist()))
    for tag, net in (("actor", ag.actor), ("critic_1", ag.critic_1), ("critic_2", ag.critic_2),
                     ("target_actor", ag.target_actor), ("target_critic_1", ag.target_critic_1),
                     ("target_critic_2", ag.target_critic_2)):
        for li, ((w, b), (rw, rb)) in enumerate(zip(net.layers(), ddpg_params_from_golden(g, n - 1, tag))):
            mx, p9999, tol, used = weight_error_report(w, rw, lr, n)
            lines.append(f"  {tag}.{li}.weight: max {mx:.2e}, 99.99 % {p9999:.2e}, allowance {tol:.2e} ({100 * used:.1f} % used)")
            if not (weights_close(w, rw, lr * eng, n, rtol=1e-5 * eng) and weights_close(b, rb, lr * eng, n, rtol=1e-5 * eng)):
                failures.append((tag, li))
    with capsys.disabled():
        print("\n" + "\n".join(lines))
    assert not failures, failures


def test_td3_samples_its_own_noise_and_saves_reference_files(tmp_path):
    import os
    from tests.helpers import her_episodes
    g = load("her_reach_small")
    from gcrl_b200 import TD3Agent
    ag = TD3Agent(10, 3, make_config(batch_size=64, max_len=100000), None, 2, 40)
    for ep in her_episodes(g):
        ag.buffer.push_episode(ep["s"], ep["a"], ep["ns"], ep["r"], ep["d"], ep["ag"], ep["fut"])
    for step in (1, 2):
        info = ag.update(step)
        assert len(info) == (8 if step % ag.ac_update_freq == 0 else 6)
        assert all(np.isfinite(float(x)) for x in info)
    ag.save_weights(str(tmp_path))
    assert sorted(os.listdir(tmp_path)) == ["actor.pth", "critic_1.pth", "critic_2.pth"]
