"""Uniform / prioritised replay on the GPU (csrc/per.cu through the C ABI) against oracle/per.py, NumPy itself
and the fixtures dumped from the unmodified reference (tests/golden/make_golden.py::per_case).

Bit-exact: the float32 priority sum, P, the float64 table, the drawn positions, the gathered rows, FIFO eviction.
float32 power goes through the platform's powf in the reference (libm or a SIMD library, depending on the host
CPU; see oracle/per.py), so: new priorities 2 ulp of float32; importance weights 4 ulp (a quotient of two powers,
each within an ulp of the correctly rounded value).
fp32 tolerance (rel 2e-5 + 1e-6 on metrics, tests.helpers.weights_close on weights): the weighted updates."""
import random
import types

import numpy as np
import pytest

from oracle import per as OP
from tests.helpers import (PER_CASES, assert_sac_actor_close, bits, ddpg_params_from_golden, load, per_initial_nets,
                           per_meta, per_pushes, per_td_position, sac_params_from_golden, ulp_diff_f32, weights_close)

pytestmark = pytest.mark.gpu


def rand_rows(rng, n, D, A):
    s = rng.standard_normal((n, D)).astype(np.float32)
    a = rng.uniform(-1, 1, (n, A)).astype(np.float32)
    r = -(rng.random(n) > 0.3).astype(np.float32)
    ns = rng.standard_normal((n, D)).astype(np.float32)
    d = (rng.random(n) < 0.1).astype(np.float32)
    return s, a, r, ns, d


def packed(s, a, r, ns, d):
    return np.concatenate([s, a, r[:, None], ns, d[:, None]], 1)


def numpy_reference_draw(prio, u, beta):
    """The reference's own NumPy calls (src/buffer.py:53-66) on a priority vector."""
    P = np.array(prio, dtype=np.float32)
    s = P.sum()
    P /= s
    cdf = P.astype(np.float64).cumsum()
    cdf /= cdf[-1]
    idx = cdf.searchsorted(u, side="right")
    w = (P.shape[0] * P[idx]) ** (-beta)
    w /= w.max()
    return s, P, cdf, idx, w


@pytest.mark.parametrize("n,cap,B", [(5, 8, 4), (130, 130, 64), (1000, 1000, 256), (4097, 5000, 256),
                                     (70001, 70001, 1024), (9000, 4097, 300), (1 << 20, 1 << 20, 4096)])
def test_prioritised_draw_is_bit_exact(n, cap, B):
    from gcrl_b200 import PERBuffer
    rng = np.random.default_rng(n)
    D, A = 6, 2
    buf = PERBuffer(cap, 0.6)
    rows = rand_rows(rng, n, D, A)
    for lo in range(0, n, 50000):                       # several commits: the ring wraps when n > cap
        buf.push_rows(*(x[lo:lo + 50000] for x in rows))
    N = min(n, cap)
    assert len(buf) == N
    live = packed(*rows)[n - N:]
    assert np.array_equal(bits(buf.rows(0, min(N, 300))), bits(live[:300]))
    prio = (rng.random(N) ** 3 * 5 + 1e-3).astype(np.float32)
    buf.set_priorities(prio)
    assert np.array_equal(bits(buf.priorities), bits(prio))
    u = rng.random(B)
    u[0] = 0.0
    for beta in (0.4, 1.0):
        s, a, r, ns, d, w, idx = buf.sample(B, beta, uniforms=u)
        want_sum, want_P, want_cdf, want_idx, want_w = numpy_reference_draw(prio, u, beta)
        got_sum, sequential = buf.last_sample_info()
        assert bits(got_sum) == bits(want_sum)
        # order-independent scan whenever every P[i] is a multiple of 2^-52 (always, if min P >= 2^-29)
        assert sequential == bool(np.any(np.mod(want_P.astype(np.float64) * 2.0 ** 52, 1.0) != 0))
        assert not sequential or N >= 100000
        got_P, got_cdf = buf.last_tables()
        assert np.array_equal(bits(got_P), bits(want_P))
        assert np.array_equal(got_cdf.view(np.uint64), want_cdf.view(np.uint64))
        assert np.array_equal(idx, want_idx)
        got = packed(s.cpu().numpy(), a.cpu().numpy(), r.cpu().numpy()[:, 0], ns.cpu().numpy(), d.cpu().numpy()[:, 0])
        assert np.array_equal(bits(got), bits(live[want_idx]))
        assert ulp_diff_f32(w.cpu().numpy()[:, 0], want_w).max() <= 4
    if N <= 5000:                                        # the restated rules agree too (pure-Python loops)
        assert np.array_equal(OP.choice_indices(OP.normalised_priorities(prio), u), want_idx)


def _priorities(kind, n, rng):
    if kind == "loguniform":                 # 15 decades: most additions of the float64 cumsum round
        return (10.0 ** rng.uniform(-14, 1, n)).astype(np.float32)
    if kind == "ties":                       # few mantissa bits, wide exponents: exact half-ulp ties everywhere
        return (rng.integers(1, 8, n) * 2.0 ** rng.integers(-40, 3, n)).astype(np.float32)
    if kind == "spikes":                     # single elements that jump several binades at once
        p = np.full(n, 1e-9, np.float32) * rng.uniform(0.5, 1.5, n).astype(np.float32)
        p[rng.integers(0, n, max(3, n // 5000))] = rng.uniform(0.5, 2.0, max(3, n // 5000))
        return p
    if kind == "zeros":                      # empty entries: repeated table values
        p = (10.0 ** rng.uniform(-12, 0, n)).astype(np.float32)
        p[rng.random(n) < 0.3] = 0.0
        p[0] = 0.0
        return p
    raise ValueError(kind)


@pytest.mark.parametrize("kind", ["loguniform", "ties", "spikes", "zeros"])
@pytest.mark.parametrize("n", [700, 40000, 300001])
def test_rounding_cumsum_is_reproduced_exactly(kind, n):
    """P[i] that is not a multiple of 2^-52 (below 2^-29 with a full mantissa): the float64 additions of the
    reference's left-to-right cumsum round -- mostly as exact ties -- and a reordered scan would differ.  The
    library must notice and track every rounding (per.cu: parity functions per 512-entry chunk)."""
    from gcrl_b200 import PERBuffer
    rng = np.random.default_rng(n)
    buf = PERBuffer(n, 0.6)
    buf.push_rows(*rand_rows(rng, n, 4, 2))
    prio = _priorities(kind, n, rng)
    buf.set_priorities(prio)
    u = rng.random(512)
    *_, idx = buf.sample(512, 0.5, uniforms=u)
    want_sum, want_P, want_cdf, want_idx, _ = numpy_reference_draw(prio, u, 0.5)
    got_sum, inexact = buf.last_sample_info()
    assert inexact == bool(np.any(np.mod(want_P.astype(np.float64) * 2.0 ** 52, 1.0) != 0))
    assert inexact or kind == "spikes"
    assert bits(got_sum) == bits(want_sum)
    got_P, got_cdf = buf.last_tables()
    assert np.array_equal(bits(got_P), bits(want_P))
    bad = np.nonzero(got_cdf.view(np.uint64) != want_cdf.view(np.uint64))[0]
    assert bad.size == 0, (bad[:5], got_cdf[bad[:5]], want_cdf[bad[:5]])
    assert np.array_equal(idx, want_idx)


def test_update_priorities_applies_in_order_and_the_last_duplicate_wins():
    import torch
    from gcrl_b200 import PERBuffer
    rng = np.random.default_rng(3)
    n, B = 500, 256
    for alpha in (0.6, 1.0, 0.0):
        buf = PERBuffer(n, alpha)
        buf.push_rows(*rand_rows(rng, n, 5, 3))
        idx = rng.integers(0, 40, B)                       # heavy duplication
        td = (rng.standard_normal((B, 1)) * 10.0 ** rng.uniform(-7, 1, (B, 1))).astype(np.float32)
        orc = OP.PERBufferOracle(n, alpha)
        for _ in range(n):
            orc.push(np.zeros(1), np.zeros(1), 0.0, np.zeros(1), 0.0)
        orc.update_priorities(idx, td)
        buf.update_priorities(idx, td)
        assert ulp_diff_f32(buf.priorities, np.array(orc.priorities, np.float32)).max() <= 2
        # device-resident TD errors + the positions of the preceding sample (the agent's path)
        u = rng.random(B)
        *_, pos = buf.sample(B, 0.4, uniforms=u)
        orc.update_priorities(pos, td)
        tdd = torch.from_numpy(td).cuda()
        buf.update_priorities_last(B, tdd.data_ptr())
        assert np.array_equal(buf.last_positions(B), pos)
        assert ulp_diff_f32(buf.priorities, np.array(orc.priorities, np.float32)).max() <= 2


def test_uniform_replay_follows_the_reference_mt_stream_and_evicts_fifo():
    from gcrl_b200 import ReplayBuffer
    rng = np.random.default_rng(0)
    D, A, cap, n = 7, 3, 333, 1000
    rows = rand_rows(rng, n, D, A)
    buf, orc = ReplayBuffer(cap), OP.ReplayBufferOracle(cap)
    with pytest.raises(AssertionError):
        buf.sample(4)
    for i in range(n):
        one = tuple(x[i] for x in rows)
        buf.push(*one)
        orc.push(*one)
        if i in (100, 332, 333, 700, 999):
            assert len(buf) == len(orc)
            random.seed(1898 + i)
            want = orc.sample(64)
            random.seed(1898 + i)
            got = buf.sample(64)
            for g_, w_ in zip(got, want):
                assert np.array_equal(bits(g_.cpu().numpy()), bits(w_))
            assert random.random() == (random.seed(1898 + i), random.sample(range(len(orc)), 64), random.random())[2]


# ---- fixtures from the unmodified reference -------------------------------------------------------------
def per_config(m, algo):
    cfg = types.SimpleNamespace(
        hidden_dim=m["H"], layer_count=m["L"], actor_lr=m["lr"], actor_lr_min=m["lr"], ac_scheduler_steps=1,
        critic_lr=m["lr"], critic_lr_min=m["lr"], cr_scheduler_steps=1, buffer_type="PER", max_len=m["max_len"],
        alpha=m["alpha"], batch_size=m["B"], gamma=m["gamma"], ac_update_freq=m["freq"], noise_std=0.2,
        noise_clamp=0.5, policy_noise=0.2, grad_clip=m["clip"], beta=m["beta"], beta_end=m["beta_end"], k_future=4,
        max_eps_len=50, tau=m["tau"])
    if algo in ("sac", "tqc"):
        cfg.alpha_lr, cfg.alpha_min, cfg.alpha_min_steps = 1e-2, 0.05, 0
    return cfg


def make_per_agent(algo, g):
    import gcrl_b200
    from gcrl_b200.agent import NET_ACTOR, NET_CRITIC, NET_CRITIC2
    m = per_meta(g)
    cfg = per_config(m, algo)
    nets = per_initial_nets(algo, g)
    if algo in ("ddpg", "td3"):
        cls = gcrl_b200.DDPG if algo == "ddpg" else gcrl_b200.TD3Agent
        ag = cls(m["D"], m["A"], cfg, None, 1, 40)
        for net, params in zip((NET_ACTOR, NET_CRITIC, NET_CRITIC2), nets):
            ag._set_layers(net, params)
        ag.update_target_network()
    else:
        from tests.test_sac_gpu import load_initial
        cls = gcrl_b200.SACAgent if algo == "sac" else gcrl_b200.TQCAgent
        ag = cls(m["D"], m["A"], cfg, None, 1, m["gstep"])
        load_initial(ag, *nets)
    assert isinstance(ag.buffer, gcrl_b200.PERBuffer)
    return ag, m


def push_range(target, g, lo, hi):
    for i in range(lo, hi):
        target.push(g["push_s"][i], g["push_a"][i], g["push_r"][i], g["push_ns"][i], g["push_d"][i])


@pytest.mark.parametrize("algo,case", PER_CASES)
def test_per_buffer_matches_reference_fixture(algo, case):
    from gcrl_b200 import PERBuffer
    g = load(f"per_{algo}_{case}")
    m = per_meta(g)
    buf = PERBuffer(m["max_len"], m["alpha"])
    for si in range(len(g["steps"])):
        push_range(buf, g, *per_pushes(g, si))
        before = g[f"s{si}_prio_before"]
        assert len(buf) == before.shape[0]
        assert ulp_diff_f32(buf.priorities, before).max() <= 2          # free-running so far
        buf.set_priorities(before)                                       # teacher-force (powf ulp)
        s, a, r, ns, d, w, idx = buf.sample(m["B"], float(g[f"s{si}_beta"]), uniforms=g[f"s{si}_u"])
        assert np.array_equal(idx, g[f"s{si}_idx"])
        assert np.array_equal(bits(buf.last_tables()[0]), bits(g[f"s{si}_P"]))
        for got, key in zip((s, a, r, ns, d), ("s", "a", "r", "ns", "d")):
            assert np.array_equal(bits(got.cpu().numpy()), bits(g[f"s{si}_batch_{key}"])), key
        assert ulp_diff_f32(w.cpu().numpy(), g[f"s{si}_w"]).max() <= 4
        buf.update_priorities(idx, g[f"s{si}_td"])
        assert ulp_diff_f32(buf.priorities, g[f"s{si}_prio_after"]).max() <= 2


@pytest.mark.parametrize("algo,case", PER_CASES)
def test_agent_update_with_prioritised_replay_matches_reference_fixture(algo, case, monkeypatch):
    """agent.update(step) end to end -- draw, weighted critic loss, per-sample TD errors, priority write-back,
    beta schedule -- against the unmodified reference agent on the same pushes, uniforms and normal draws."""
    import torch
    g = load(f"per_{algo}_{case}")
    ag, m = make_per_agent(algo, g)
    B, lr = m["B"], m["lr"]
    steps = [int(x) for x in g["steps"]]
    for si, step in enumerate(steps):
        push_range(ag, g, *per_pushes(g, si))
        assert ag.is_buffer_filled()
        ag.buffer.set_priorities(g[f"s{si}_prio_before"])               # teacher-force: see module docstring
        assert ag.beta == pytest.approx(float(g[f"s{si}_beta"]), rel=0, abs=0)
        monkeypatch.setattr(np.random, "random_sample", lambda n, _u=g[f"s{si}_u"]: _u.copy())
        normals = [torch.from_numpy(g[f"s{si}_normal{j}"]).cuda() for j in range(int(g[f"s{si}_n_normal"]))]
        if algo == "ddpg":
            info = ag.update(step)
        elif algo == "td3":
            info = ag.update(step, noise=normals[0])
        else:
            info = ag.update(step, eps_next=normals[0], eps_cur=normals[1] if len(normals) > 1 else None)
        monkeypatch.undo()
        ref = g[f"s{si}_info"]
        assert len(info) == len(ref)
        tdp = per_td_position(algo, len(info))
        td = info[tdp]
        assert isinstance(td, np.ndarray) and td.shape == (B, 1) and td.dtype == np.float32
        assert np.array_equal(ag.buffer.last_positions(B), g[f"s{si}_idx"])
        np.testing.assert_allclose(td, g[f"s{si}_td"], rtol=5e-5, atol=5e-6)
        flat = [float(np.mean(x)) if i == tdp else float(x) for i, x in enumerate(info)]
        np.testing.assert_allclose(np.array(flat), ref, rtol=5e-5, atol=2e-6)
        # the priorities written back are the rule of src/buffer.py:89 applied to the TD errors returned
        orc = OP.PERBufferOracle(m["max_len"], m["alpha"])
        orc.priorities.extend(g[f"s{si}_prio_before"].tolist())
        orc.update_priorities(g[f"s{si}_idx"], td)
        assert ulp_diff_f32(ag.buffer.priorities, np.array(orc.priorities, np.float32)).max() <= 2
        np.testing.assert_allclose(ag.buffer.priorities, g[f"s{si}_prio_after"], rtol=1e-4, atol=1e-5)
    si, n = len(steps) - 1, len(steps)
    if f"s{si}_actor.base_net.0.weight" not in g.files:
        return
    if algo in ("ddpg", "td3"):
        views = {"actor": ag.actor, "target_actor": ag.target_actor}
        if algo == "ddpg":
            views.update(critic=ag.critic, target_critic=ag.target_critic)
        else:
            views.update(critic_1=ag.critic_1, critic_2=ag.critic_2, target_critic_1=ag.target_critic_1,
                         target_critic_2=ag.target_critic_2)
        for tag, view in views.items():
            for (w, b), (rw, rb) in zip(view.layers(), ddpg_params_from_golden(g, si, tag)):
                assert weights_close(w, rw, lr, n) and weights_close(b, rb, lr, n), tag
    else:
        from tests.test_sac_gpu import actor_params
        params, _ = actor_params(ag)
        assert_sac_actor_close(params, sac_params_from_golden(g, si, "actor")["params"], lr, n)
        for i in (0, ag.N_CRITICS - 1):
            tag = f"critic_{i + 1}" if algo == "sac" else f"critic_{i}"
            for (w, b), (rw, rb) in zip(ag._critic_views[i].layers(), sac_params_from_golden(g, si, tag)):
                assert weights_close(w, rw, lr, n) and weights_close(b, rb, lr, n), tag


def test_agent_update_with_uniform_replay_equals_the_explicit_batch_update():
    """buffer_type "REPLAY" (src/agent.py:69-70): update(step) == sample with random.sample's stream + update on
    that batch, bit for bit."""
    import gcrl_b200
    from gcrl_b200.agent import NET_ACTOR, NET_CRITIC
    from oracle import ddpg as OD
    from tests.test_ddpg_gpu import make_config
    D, A, H, L, B = 13, 3, 64, 3, 96
    rng = np.random.default_rng(4)
    rows = rand_rows(rng, 400, D, A)
    agents = []
    for _ in range(2):
        ag = gcrl_b200.DDPG(D, A, make_config(hidden_dim=H, layer_count=L, batch_size=B, buffer_type="REPLAY", max_len=300),
                            None, 1, 40)
        r2 = np.random.default_rng(9)
        ag._set_layers(NET_ACTOR, OD.init_mlp(r2, D, H, A, L))
        ag._set_layers(NET_CRITIC, OD.init_mlp(r2, D + A, H, 1, L))
        ag.update_target_network()
        assert not ag.is_buffer_filled()
        for i in range(400):
            ag.push(*(x[i] for x in rows))
        assert len(ag.buffer) == 300 and ag.is_buffer_filled()
        agents.append(ag)
    a1, a2 = agents
    for step in (39, 40, 41):
        random.seed(step)
        i1 = a1.update(step)
        random.seed(step)
        i2 = a2.update(step, batch=a2.buffer.sample(B))
        assert [float(x) for x in i1] == [float(x) for x in i2]
    for (w, b), (w2, b2) in zip(a1.actor.layers() + a1.critic.layers(), a2.actor.layers() + a2.critic.layers()):
        assert np.array_equal(w, w2) and np.array_equal(b, b2)


@pytest.mark.parametrize("btype", ["PER", "REPLAY"])
def test_true_resume_with_replay_buffers_is_bit_identical(tmp_path, btype):
    """save_checkpoint / load_checkpoint carry the ring (deque order), the priorities and both host generators
    (random for ReplayBuffer, numpy.random for PERBuffer): the resumed run draws the same positions and makes
    bit-identical updates."""
    import torch
    import gcrl_b200
    from tests.test_ddpg_gpu import make_config
    D, A, H, L, B = 13, 3, 64, 3, 96
    rng = np.random.default_rng(12)
    rows = rand_rows(rng, 900, D, A)
    cfg = make_config(hidden_dim=H, layer_count=L, batch_size=B, buffer_type=btype, max_len=500, alpha=0.6, beta=0.4,
                      beta_end=50)
    torch.manual_seed(5)
    a1 = gcrl_b200.DDPG(D, A, cfg, None, 1, 40)
    for i in range(700):
        a1.push(*(x[i] for x in rows))
    random.seed(3)
    np.random.seed(3)
    for i in range(4):
        a1.update(37 + i)
    a1.save_checkpoint(str(tmp_path / "ck"))
    def cont(ag):
        out = []
        for i in range(4):
            for j in range(700 + 50 * i, 750 + 50 * i):
                ag.push(*(x[j] for x in rows))
            out.append(ag.update(41 + i))
        return out
    ref = cont(a1)
    random.seed(77)
    np.random.seed(77)
    torch.manual_seed(99)
    a2 = gcrl_b200.DDPG(D, A, cfg, None, 1, 40)
    a2.load_checkpoint(str(tmp_path / "ck"))
    assert len(a2.buffer) == 500
    got = cont(a2)
    for r, g_ in zip(ref, got):
        assert len(r) == len(g_)
        for x, y in zip(r, g_):
            assert np.array_equal(np.asarray(x), np.asarray(y))
    assert np.array_equal(bits(a1.buffer.rows()), bits(a2.buffer.rows()))
    if btype == "PER":
        assert np.array_equal(bits(a1.buffer.priorities), bits(a2.buffer.priorities))
        assert a1.beta == a2.beta
