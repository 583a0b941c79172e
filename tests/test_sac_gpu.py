"""GPU parity: the SAC / TQC update (csrc/sac.cu through gcrl_sac_*) versus fixtures dumped from the
unmodified reference ``SACAgent.update`` / ``TQCAgent.update`` (src/agent.py:659-699, :1062-1100)
with the recorded ``Normal.rsample`` noise, and versus the NumPy oracle on ragged shapes.

Tolerance: metrics rel 2e-5 + abs 1e-6; weights tests.helpers.weights_close /
assert_sac_actor_close (pre-BatchNorm Linear biases only to Adam's hard bound -- their true
gradient is zero, see the helper); BatchNorm running variance rel 1e-5."""
import ctypes as C
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


def _rand_batch(rng, B, D, A):
    import torch
    s = rng.standard_normal((B, D)).astype(np.float32)
    arrs = (s, rng.uniform(-1, 1, (B, A)).astype(np.float32), -(rng.random((B, 1)) > 0.3).astype(np.float32),
            (s + 0.1 * rng.standard_normal((B, D))).astype(np.float32), (rng.random((B, 1)) < 0.1).astype(np.float32),
            rng.standard_normal((B, A)).astype(np.float32), rng.standard_normal((B, A)).astype(np.float32))
    return tuple(torch.from_numpy(x).cuda() for x in arrs)


def _make(algo, D, A, H, L, B, seed=5):
    from gcrl_b200 import SACAgent, TQCAgent
    from oracle import ddpg as OD
    from oracle import sac as OS
    rng = np.random.default_rng(seed)
    cfg = sac_config(hidden_dim=H, layer_count=L, batch_size=B, grad_clip=0.5, tau=0.05, alpha_min_steps=0, alpha_lr=1e-2)
    actor0, stats0 = OS.init_sac_actor(rng, D, H, A, L, head_scale=0.1, log_std_bias=-1.0)
    critics0 = [OD.init_mlp(rng, D + A, H, 1, L) for _ in range(2 if algo == "sac" else 5)]
    ag = (SACAgent if algo == "sac" else TQCAgent)(D, A, cfg, None, 1, 2)
    load_initial(ag, actor0, stats0, critics0)
    return ag


@pytest.mark.parametrize("sync_bn", [False, True])
@pytest.mark.parametrize("algo", ["sac", "tqc"])
def test_world1_phases_equal_whole_update_bitwise(algo, sync_bn):
    """World of one: the four data-parallel phases (local BatchNorm statistics) and the sync-BN segment chain
    (statistics merged over the ranks' slots) both reproduce the whole update bit for bit."""
    D, A, H, L, B = 22, 3, 64, 3, 128
    whole, phased = _make(algo, D, A, H, L, B), _make(algo, D, A, H, L, B)
    phased.enable_data_parallel(allreduce_mean=lambda t: t, sync_bn=sync_bn, world=1, rank=0,
                                allgather=lambda t: t)                 # world of one: identity
    rng = np.random.default_rng(0)
    for step in (1, 2, 3):
        *batch, e1, e2 = _rand_batch(rng, B, D, A)
        i1 = whole.update(step, batch=tuple(batch), eps_next=e1, eps_cur=e2)
        i2 = phased.update(step, batch=tuple(batch), eps_next=e1, eps_cur=e2)
        assert [float(x) for x in i1] == [float(x) for x in i2]
    p1, s1 = actor_params(whole)
    p2, s2 = actor_params(phased)
    for (w, b), (w2, b2) in zip(p1 + s1, p2 + s2):
        assert np.array_equal(w, w2) and np.array_equal(b, b2)
    for v1, v2 in zip(whole._critic_views + whole._target_views, phased._critic_views + phased._target_views):
        for (w, b), (w2, b2) in zip(v1.layers(), v2.layers()):
            assert np.array_equal(w, w2) and np.array_equal(b, b2)
    assert whole.get_log_alpha() == phased.get_log_alpha()


def test_tqc_two_emulated_ranks_stay_replicated_and_track_the_averaged_gradient():
    """Two agents on one GPU play ranks 0/1: phases in lock step, the buffers of gcrl_sac_dp_buffer averaged
    between them exactly as the NCCL all-reduce does.  Replicas must stay bit-identical (weights, Adam
    trajectories, log_alpha, BatchNorm running statistics); BatchNorm normalises with LOCAL batch statistics
    (DDP's default), so the result is close to, not equal to, one rank on the concatenated batch."""
    import torch
    from gcrl_b200._lib import check, lib, vp
    D, A, H, L, B = 22, 3, 64, 2, 256
    ranks = [_make("tqc", D, A, H, L, B), _make("tqc", D, A, H, L, B)]
    single = _make("tqc", D, A, H, L, 2 * B)
    rng = np.random.default_rng(1)
    st = vp(torch.cuda.current_stream().cuda_stream)

    def average(which):
        g = [ag.grad_tensor(which) for ag in ranks]
        mean = (g[0] + g[1]) / 2
        g[0].copy_(mean)
        g[1].copy_(mean)
    for step in (1, 2, 3):
        data = [_rand_batch(rng, B, D, A) for _ in ranks]
        flags = 1 | 2 | 4
        for phase in range(4):
            for ag, b in zip(ranks, data):
                check(lib.gcrl_sac_update_phase(ag._h, phase, None, B, None, *(vp(t.data_ptr()) for t in b[:5]),
                                                vp(b[5].data_ptr()), vp(b[6].data_ptr()), 1e-3, 1e-3, flags, st))
            if phase == 0:
                average(1)
            elif phase == 2:
                average(0)
        average(2)
        cat = [torch.cat([x, y]) for x, y in zip(*data)]
        info = single.update(step, batch=tuple(cat[:5]), eps_next=cat[5], eps_cur=cat[6])
        m = []
        for ag in ranks:
            buf = (C.c_float * 12)()
            check(lib.gcrl_sac_read_metrics(ag._h, flags, C.cast(buf, vp), st))
            m.append(np.array(list(buf)))
        got = (m[0] + m[1]) / 2
        want = np.array([float(x) for x in info])
        np.testing.assert_allclose(got[[0, 3, 4]], want[[0, 3, 4]], rtol=2e-2)       # losses / td / q: means of means
        assert m[0][9] == m[1][9] and m[0][10] == m[1][10]                           # alpha, log_alpha replicated
    p0, s0 = actor_params(ranks[0])
    p1, s1 = actor_params(ranks[1])
    ps, _ = actor_params(single)
    for (w, b), (w2, b2) in zip(p0 + s0, p1 + s1):
        assert np.array_equal(w, w2) and np.array_equal(b, b2)
    for v0, v1 in zip(ranks[0]._critic_views + ranks[0]._target_views, ranks[1]._critic_views + ranks[1]._target_views):
        for (w, b), (w2, b2) in zip(v0.layers(), v1.layers()):
            assert np.array_equal(w, w2) and np.array_equal(b, b2)
    for (w, _), (ws, _) in zip(ranks[0]._critic_views[0].layers(), single._critic_views[0].layers()):
        assert np.max(np.abs(w - ws)) <= 2.0 * 1e-3 * 3                              # Adam's hard bound


@pytest.mark.parametrize("algo,world", [("sac", 2), ("tqc", 2), ("tqc", 4)])
def test_sync_bn_ranks_equal_one_rank_on_the_concatenated_batch(algo, world):
    """with BatchNorm statistics taken over the global batch, `world` ranks on B rows each must equal ONE
    rank on the concatenated world * B rows -- metrics and weights at the DDPG tolerance (tests/helpers.py), not the
    2e-2 of the local-statistics mode.  The ranks are agents on one GPU driven in lock step through
    gcrl_sac_update_segment; the collectives (slot all-gather, gradient averages) are done by hand exactly as NCCL does
    them.  Replicas must stay bit-identical, running statistics included, without averaging them."""
    import torch
    from gcrl_b200._lib import check, lib, vp
    D, A, H, L, B = 22, 3, 64, 3, 128
    ranks = [_make(algo, D, A, H, L, B) for _ in range(world)]
    single = _make(algo, D, A, H, L, world * B)
    for r, ag in enumerate(ranks):
        check(lib.gcrl_sac_set_sync_bn(ag._h, world, r))
    slots = [ag.grad_tensor(4).view(world, -1) for ag in ranks]
    rng = np.random.default_rng(2)
    st = vp(torch.cuda.current_stream().cuda_stream)
    nsteps, lr = 4, 1e-3
    counts = {1: 0, 2: 0, 3: 0}
    for step in range(1, nsteps + 1):
        data = [_rand_batch(rng, B, D, A) for _ in ranks]
        flags = (1 if step % 2 == 0 else 0) | 2 | (4 if step % 2 == 0 else 0)
        seg = 0
        while True:
            colls = []
            for ag, b in zip(ranks, data):
                coll = C.c_int(-1)
                check(lib.gcrl_sac_update_segment(ag._h, seg, None, B, None, *(vp(t.data_ptr()) for t in b[:5]),
                                                  vp(b[5].data_ptr()), vp(b[6].data_ptr()), lr, lr, flags,
                                                  C.byref(coll), st))
                colls.append(coll.value)
            assert len(set(colls)) == 1
            if colls[0] == 0:
                break
            counts[colls[0]] += 1
            if colls[0] == 1:                                   # all-gather: rank r contributes row r
                for r in range(world):
                    for q in range(world):
                        if q != r:
                            slots[q][r].copy_(slots[r][r])
            else:
                g = [ag.grad_tensor(1 if colls[0] == 2 else 0) for ag in ranks]
                mean = sum(g[1:], g[0].clone()) / world
                for t in g:
                    t.copy_(mean)
            seg += 1
        cat = [torch.cat(x) for x in zip(*data)]
        info = _update_with_flags(single, step, cat, flags, lr)
        m = []
        for ag in ranks:
            buf = (C.c_float * 12)()
            check(lib.gcrl_sac_read_metrics(ag._h, flags, C.cast(buf, vp), st))
            m.append(np.array(list(buf)))
        got = sum(m) / world
        np.testing.assert_allclose(got[:5], info[:5], rtol=2e-5, atol=1e-6)          # losses / td / q
        assert all(np.array_equal(m[0][9:11], x[9:11]) for x in m[1:])              # alpha, log_alpha replicated
    # per update: L gathers for the target-policy forward (+ 2 L on actor steps), one critic average, one actor average
    assert counts == {1: L * nsteps + 2 * L * (nsteps // 2), 2: nsteps, 3: nsteps // 2}
    p0, s0 = actor_params(ranks[0])
    for ag in ranks[1:]:
        p1, s1 = actor_params(ag)
        for (w, b), (w2, b2) in zip(p0 + s0, p1 + s1):
            assert np.array_equal(w, w2) and np.array_equal(b, b2)
        for v0, v1 in zip(ranks[0]._critic_views + ranks[0]._target_views, ag._critic_views + ag._target_views):
            for (w, b), (w2, b2) in zip(v0.layers(), v1.layers()):
                assert np.array_equal(w, w2) and np.array_equal(b, b2)
    ps, ss = actor_params(single)
    assert_sac_actor_close(p0, ps, lr, nsteps)      # pre-BatchNorm biases (zero true gradient): Adam's hard bound
    assert_running_stats_close(p0, s0, ps, ss)
    for i, (v0, vs) in enumerate(zip(ranks[0]._critic_views + ranks[0]._target_views,
                                     single._critic_views + single._target_views)):
        for k, ((w, b), (ws, bs)) in enumerate(zip(v0.layers(), vs.layers())):
            assert weights_close(w, ws, lr, nsteps) and weights_close(b, bs, lr, nsteps), f"critic {i} layer {k}"
    assert abs(ranks[0].get_log_alpha() - single.get_log_alpha()) <= 1e-6
    # a sync-BN agent refuses the whole-update entry points instead of silently using local statistics
    with pytest.raises(ValueError):
        _update_with_flags(ranks[0], 9, [x[:B] for x in cat], flags, lr)


def assert_running_stats_close(params, stats, ref_params, ref_stats):
    """BatchNorm running statistics of two runs that should agree.  The running mean of layer l is an average of
    batch means of z = W x + b, so it inherits the drift of the pre-BatchNorm bias b one to one -- and that bias has
    an exactly-zero true gradient, i.e. AdamW moves it by rounding noise (assert_sac_actor_close holds it to Adam's
    hard bound only; it cannot change the train-mode output).  Allowance: 3 x the observed bias difference + 2e-6;
    the running variance does not see the bias and is held to rel 2e-5."""
    for k, ((rm, rv), (rms, rvs)) in enumerate(zip(stats, ref_stats)):
        db = float(np.max(np.abs(np.asarray(params[2 * k][1], np.float64) - ref_params[2 * k][1])))
        np.testing.assert_allclose(rm, rms, rtol=1e-5, atol=3.0 * db + 2e-6, err_msg=f"running_mean {k}")
        np.testing.assert_allclose(rv, rvs, rtol=2e-5, atol=1e-7, err_msg=f"running_var {k}")


def _update_with_flags(ag, step, cat, flags, lr):
    """One whole update through the C ABI with explicit flags / learning rates (the Python update() derives them from
    the step number and the schedulers)."""
    import torch
    from gcrl_b200._lib import check, lib, vp
    buf = (C.c_float * 12)()
    st = vp(torch.cuda.current_stream().cuda_stream)
    check(lib.gcrl_sac_update_batch(ag._h, cat[0].shape[0], *(vp(t.data_ptr()) for t in cat[:5]),
                                    vp(cat[5].data_ptr()), vp(cat[6].data_ptr()), lr, lr, flags, C.cast(buf, vp), st))
    return np.array(list(buf))


@pytest.mark.parametrize("algo", ["sac", "tqc"])
def test_reset_redraws_linears_only_and_handle_types_do_not_mix(algo):
    """reset() as SACAgent.reset / TQCAgent.reset (src/agent.py:755-769, :1161-1170): xavier re-draw of the nn.Linear
    layers of the actor, the critics AND (independently) the target critics; BatchNorm parameters and running
    statistics survive; log_alpha back to 0.  The DDPG-only services of the shared base class refuse loudly, and
    the C side rejects a handle of the other agent family instead of reinterpreting it."""
    import torch
    from gcrl_b200 import SACAgent, TQCAgent
    from gcrl_b200._lib import GcrlError, check, lib
    cls = SACAgent if algo == "sac" else TQCAgent
    D, A, H, L, B = 22, 3, 64, 2, 32
    torch.manual_seed(5)
    ag = cls(D, A, sac_config(hidden_dim=H, layer_count=L, batch_size=B), None, 1, 40)
    rng = np.random.default_rng(0)
    gam, bet, rm, rv = (rng.standard_normal(H).astype(np.float32) for _ in range(4))
    ag.actor.set_bn(0, gam, bet, rm, np.abs(rv) + 0.5)
    ag.set_log_alpha(-0.7)
    before_lin = ag.actor.linear(0)[0].copy()
    before_c = ag._critic_views[0].layers()[0][0].copy()
    ag.reset()
    g2, b2, rm2, rv2 = ag.actor.bn(0)
    assert np.array_equal(g2, gam) and np.array_equal(b2, bet) and np.array_equal(rm2, rm) and np.array_equal(rv2, np.abs(rv) + 0.5)
    assert not np.array_equal(ag.actor.linear(0)[0], before_lin)
    assert not np.array_equal(ag._critic_views[0].layers()[0][0], before_c)
    for c, t in zip(ag._critic_views, ag._target_views):          # independent draws, not copies
        assert not np.array_equal(c.layers()[0][0], t.layers()[0][0])
        assert np.all(c.layers()[0][1] == np.float32(0.01)) and np.all(t.layers()[0][1] == np.float32(0.01))
    bound = np.sqrt(6.0 / (H + D))
    assert np.abs(ag.actor.linear(0)[0]).max() <= bound + 1e-6
    assert float(ag.get_log_alpha()) == 0.0
    for name, call in (("state_dict", lambda: ag.state_dict()), ("save_checkpoint", lambda: ag.save_checkpoint("/tmp/x")),
                       ("load_checkpoint", lambda: ag.load_checkpoint("/tmp/x")), ("q_values", lambda: ag.q_values(None, None)),
                       ("enable_peer_data_parallel", lambda: ag.enable_peer_data_parallel())):
        with pytest.raises(NotImplementedError, match=name):
            call()
    # a SAC / TQC handle is not a gcrl_agent, and vice versa
    with pytest.raises(ValueError, match="handle is not"):
        check(lib.gcrl_agent_hard_update(ag._h, None))
    from gcrl_b200 import DDPG
    dd = DDPG(D, A, make_config(hidden_dim=H, layer_count=L, batch_size=B), None, 1, 40)
    with pytest.raises(ValueError, match="handle is not"):
        check(lib.gcrl_sac_hard_update(dd._h, None))
    assert lib.gcrl_agent_num_layers(ag._h, 0) == -1
    # select_action: any number of envs (more than max_batch), exploration noise from torch's generator only
    np.random.seed(3)
    state = np.random.get_state()[1].copy()
    obs = rng.standard_normal((3 * B + 5, D)).astype(np.float32)
    torch.manual_seed(9)
    a1 = ag.select_action(obs)
    torch.manual_seed(9)
    a2 = ag.select_action(obs)
    assert a1.shape == (3 * B + 5, A) and np.array_equal(a1, a2) and np.all(np.abs(a1) <= 1.0)
    assert np.array_equal(np.random.get_state()[1], state), "NumPy's global stream must not be consumed"
