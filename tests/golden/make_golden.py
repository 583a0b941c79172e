#!/usr/bin/env python
"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (needs /root/reference, which does not exist on
the GPU box):

    python tests/golden/make_golden.py

It imports the reference's own ``src.buffer`` / ``src.agent`` / ``src.utils`` /
``src.model`` behind a stub ``gymnasium`` module (``src/utils.py:5`` imports it
at module top; the real package is not installed), drives them with seeded
synthetic inputs, records every Mersenne-Twister draw the hot path consumes
(``random.randint`` at src/buffer.py:153, ``random.sample`` at src/buffer.py:124)
and stores inputs + draws + outputs as small ``.npz`` files.  The parity tests
replay those inputs/draws through the oracle (CPU) and the CUDA path (GPU).

``compute_reward`` is the one piece that cannot come from the reference tree
(panda-gym is un-vendored): the generator binds ``panda_reward`` below, the
published sparse rule written with the same NumPy calls panda-gym uses
(``np.linalg.norm(a - b, axis=-1)`` then ``-np.array(d > 0.05, float32)``).
"""
import os
import random
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("GCRL_REFERENCE", "/root/reference")


def import_reference():
    g = types.ModuleType("gymnasium")

    class _W:
        def __init__(self, env=None):
            self.env = env

    g.Wrapper = _W
    g.ObservationWrapper = _W
    g.vector = types.SimpleNamespace(AsyncVectorEnv=object,
                                     AutoresetMode=types.SimpleNamespace(NEXT_STEP=0))
    g.spaces = types.SimpleNamespace(Dict=dict, Box=object)
    sys.modules.setdefault("gymnasium", g)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import src.agent as agent
    import src.buffer as buffer
    import src.model as model
    import src.utils as utils
    return agent, buffer, model, utils


def panda_reward(achieved_goal, desired_goal, info):
    d = np.linalg.norm(achieved_goal - desired_goal, axis=-1)
    return -np.array(d > 0.05, dtype=np.float32)


# ----------------------------------------------------------------------------
# synthetic episodes (SURVEY 8d recipe, small)
# ----------------------------------------------------------------------------
def synth_episode(rng, T, O, G, A, terminated_last=False):
    """Rows exactly as env.py:163-224 hands them to push_her (no normalisation):
    state = obs_t || dg, next_state = obs_{t+1} || dg, ag = achieved goal of the
    NEXT observation, reward = sparse reward of the next observation."""
    obs = rng.standard_normal((T + 1, O)).astype(np.float32)
    ag = np.empty((T + 1, G), np.float32)
    ag[0] = rng.uniform(-0.15, 0.15, G)
    for t in range(T):
        step = rng.normal(0, 0.02, G) * (rng.random() > 0.3)
        ag[t + 1] = ag[t] + step.astype(np.float32)
    dg = rng.uniform(-0.15, 0.15, G).astype(np.float32)
    act = rng.uniform(-1, 1, (T, A)).astype(np.float32)
    s = np.concatenate([obs[:-1], np.repeat(dg[None], T, 0)], -1)
    ns = np.concatenate([obs[1:], np.repeat(dg[None], T, 0)], -1)
    rew = np.array([panda_reward(ag[t + 1], dg, {}) for t in range(T)], dtype=np.float64)
    done = np.zeros(T, bool)
    done[-1] = terminated_last
    return dict(s=s, a=act, ns=ns, r=rew, d=done, dg=np.repeat(dg[None], T, 0), ag=ag[1:])


def her_case(buffer_mod, name, *, O, G, A, k, lens, max_mem_len, nenvs, seed, batches):
    import torch
    rng = np.random.default_rng(seed)
    random.seed(1898 + seed)               # reference default seed, src/main.py:79
    buf = buffer_mod.HERBuffer(max_mem_len, 50, nenvs, k_future=k)
    buf.compute_reward = panda_reward
    draws = []
    real_randint = random.randint

    def rec_randint(a, b):
        v = real_randint(a, b)
        draws.append(v)
        return v

    out = {}
    eps = [synth_episode(rng, T, O, G, A, terminated_last=(T < 50)) for T in lens]
    # interleave the per-env staging like env.py:192: round-robin over envs
    cursors = [(i, 0) for i in range(min(nenvs, len(eps)))]
    nxt = len(cursors)
    slot_of = {i: i for i in range(len(cursors))}   # env slot -> episode id
    order = []                                      # commit order of episodes
    buffer_mod.random.randint = rec_randint
    try:
        active = dict(slot_of)
        pos = {e: 0 for e in active.values()}
        while active:
            for slot in sorted(active):
                e = active[slot]
                ep = eps[e]
                t = pos[e]
                n0 = len(draws)
                buf.push(slot, torch.from_numpy(ep["s"][t]), ep["a"][t],
                         torch.from_numpy(ep["ns"][t]), ep["r"][t], ep["d"][t],
                         ep["dg"][t], ep["ag"][t])
                pos[e] += 1
                if pos[e] == ep["s"].shape[0]:
                    T = pos[e]
                    fut = np.zeros((T, k), np.uint8)
                    d = draws[n0:]
                    assert len(d) == (T - 1) * k
                    if T > 1:
                        fut[:T - 1] = np.array(d, np.uint8).reshape(T - 1, k)
                    order.append(e)
                    out[f"ep{len(order) - 1}_fut"] = fut
                    for key, val in ep.items():
                        out[f"ep{len(order) - 1}_{key}"] = val
                    if nxt < len(eps):
                        active[slot] = nxt
                        pos[nxt] = 0
                        nxt += 1
                    else:
                        del active[slot]
    finally:
        buffer_mod.random.randint = real_randint
    out["n_episodes"] = np.int64(len(order))
    out["len"] = np.int64(len(buf))
    out["meta"] = np.array([O, G, A, k, max_mem_len, nenvs], np.int64)
    # full dump of the live deque, entry by entry
    S, Aa, NS, R, Dn, _, _ = zip(*list(buf.buffer))
    out["dump_s"] = np.array(S, np.float32)
    out["dump_a"] = np.array(Aa, np.float32)
    out["dump_ns"] = np.array(NS, np.float32)
    out["dump_r"] = np.array([np.float32(x) for x in R], np.float32)
    out["dump_d"] = np.array([np.float32(bool(x)) for x in Dn], np.float32)
    # reference sample(): recover the index stream by replaying the MT state
    for bi, B in enumerate(batches):
        st = random.getstate()
        s, a, r, ns, d = buf.sample(B)
        random.setstate(st)
        idx = random.sample(range(len(buf)), B)
        assert np.array_equal(out["dump_s"][idx], s.numpy())
        out[f"b{bi}_idx"] = np.array(idx, np.int64)
        out[f"b{bi}_s"] = s.numpy()
        out[f"b{bi}_a"] = a.numpy()
        out[f"b{bi}_r"] = r.numpy()
        out[f"b{bi}_ns"] = ns.numpy()
        out[f"b{bi}_d"] = d.numpy()
    out["n_batches"] = np.int64(len(batches))
    np.savez_compressed(os.path.join(HERE, f"her_{name}.npz"), **out)
    print(f"her_{name}: {len(order)} episodes, {len(buf)} entries")


def reward_case():
    """Known-answer edge cases + random cases for the restated reward rule."""
    rng = np.random.default_rng(7)
    a = rng.uniform(-0.2, 0.2, (4096, 3)).astype(np.float32)
    b = (a + rng.normal(0, 0.03, (4096, 3))).astype(np.float32)
    thr = np.float32(0.05)
    edge = [thr, np.nextafter(thr, np.float32(1)), np.nextafter(thr, np.float32(0)),
            np.float32(0.0)]
    ea = np.zeros((len(edge) * 3, 3), np.float32)
    eb = np.zeros_like(ea)
    for i, e in enumerate(edge):
        for ax in range(3):
            eb[i * 3 + ax, ax] = e
    a = np.concatenate([a, ea])
    b = np.concatenate([b, eb])
    r = panda_reward(a, b, {})
    np.savez_compressed(os.path.join(HERE, "reward_kat.npz"), a=a, b=b, r=r)
    print("reward_kat:", r.shape, "success frac", float(np.mean(np.signbit(r) & (r == 0))))


def normalizer_case(utils_mod):
    rng = np.random.default_rng(11)
    out = {}
    for tag, dim, dtype in (("obs", 19, np.float64), ("dg", 3, np.float32)):
        nz = utils_mod.RunningNormalizer(size=dim)
        for i in range(6):
            n = int(rng.integers(1, 40))
            x = (rng.standard_normal((n, dim)) * rng.uniform(0.1, 3) + rng.uniform(-2, 2)).astype(dtype)
            nz.update(x)
            out[f"{tag}_x{i}"] = x
            out[f"{tag}_mean{i}"] = np.array(nz.mean, np.float64)
            out[f"{tag}_var{i}"] = np.array(nz.var, np.float64)
            out[f"{tag}_count{i}"] = np.float64(nz.count)
        q = (rng.standard_normal((33, dim)) * 8).astype(dtype)
        out[f"{tag}_q"] = q
        out[f"{tag}_qn"] = nz.normalize(q)
    np.savez_compressed(os.path.join(HERE, "normalizer.npz"), **out)
    print("normalizer: ok")


def flat_params(module):
    return {k: v.detach().cpu().numpy().copy() for k, v in module.state_dict().items()}


def ddpg_case(agent_mod, utils_mod, name, *, D, A, H, L, B, steps, seed, gamma=0.98,
              tau=0.05, grad_clip=10.0, actor_lr=1e-3, critic_lr=1e-3, actor_lr_min=None,
              critic_lr_min=None, ac_T=1, cr_T=1, ac_update_freq=1, store_weights=True, store_batches=True):
    """store_batches=False (the large-batch cases): the batches are NOT written to the fixture; the test
    regenerates them from `seed` with the same NumPy generator calls, in the same order, as below."""
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import ddpg as O  # only for the seeded weight initialiser
    torch.set_num_threads(1)
    cfg = utils_mod.BaseAgentConfig(
        hidden_dim=H, layer_count=L, actor_lr=actor_lr,
        actor_lr_min=actor_lr if actor_lr_min is None else actor_lr_min,
        ac_scheduler_steps=ac_T, critic_lr=critic_lr,
        critic_lr_min=critic_lr if critic_lr_min is None else critic_lr_min,
        cr_scheduler_steps=cr_T, buffer_type="HER", max_len=1000, alpha=1.0, batch_size=B,
        gamma=gamma, ac_update_freq=ac_update_freq, noise_std=0.2, noise_clamp=0.5,
        policy_noise=0.0, grad_clip=grad_clip, beta=1.0, beta_end=1, k_future=4,
        max_eps_len=50, tau=tau)
    ag = agent_mod.DDPG(obs_dim=D, ac_dim=A, config=cfg, weights=None, nenvs=1, gradient_step=40)
    ag.device = "cpu"
    rng = np.random.default_rng(seed)
    actor0 = O.init_mlp(rng, D, H, A, L)
    critic0 = O.init_mlp(rng, D + A, H, 1, L)
    with torch.no_grad():
        for net, pref, params in ((ag.actor, "base_net", actor0), (ag.critic, "net", critic0)):
            sd = net.state_dict()
            for i, (w, b) in enumerate(params):
                sd[f"{pref}.{2 * i}.weight"].copy_(torch.from_numpy(w))
                sd[f"{pref}.{2 * i}.bias"].copy_(torch.from_numpy(b))
    ag.update_target_network()
    out = {"meta": np.array([D, A, H, L, B, seed, ac_T, cr_T, ac_update_freq], np.int64),
           "hp": np.array([gamma, tau, grad_clip, actor_lr, critic_lr,
                           actor_lr if actor_lr_min is None else actor_lr_min,
                           critic_lr if critic_lr_min is None else critic_lr_min], np.float64),
           "steps": np.array(steps, np.int64)}
    for si, step in enumerate(steps):
        s = rng.standard_normal((B, D)).astype(np.float32)
        ns = (s + 0.1 * rng.standard_normal((B, D))).astype(np.float32)
        a = rng.uniform(-1, 1, (B, A)).astype(np.float32)
        r = -(rng.random((B, 1)) > 0.3).astype(np.float32)          # -1.0 / -0.0
        d = (rng.random((B, 1)) < 0.1).astype(np.float32)
        batch = tuple(torch.from_numpy(x) for x in (s, a, r, ns, d))
        ag.buffer.sample = lambda bs, _b=batch: _b                   # feed explicit batch
        info = ag.update(step)
        if store_batches:
            out[f"s{si}_batch_s"], out[f"s{si}_batch_a"], out[f"s{si}_batch_r"] = s, a, r
            out[f"s{si}_batch_ns"], out[f"s{si}_batch_d"] = ns, d
        else:                                   # a checksum pins the regenerated stream
            out[f"s{si}_batch_sum"] = np.array([x.astype(np.float64).sum() for x in (s, a, r, ns, d)])
        out[f"s{si}_info"] = np.array([float(x) for x in info], np.float64)
        out[f"s{si}_lr"] = np.array([ag.critic_opt.param_groups[0]["lr"],
                                     ag.actor_opt.param_groups[0]["lr"]], np.float64)
        if store_weights or si == len(steps) - 1:
            for tag, net in (("actor", ag.actor), ("critic", ag.critic),
                             ("target_actor", ag.target_actor), ("target_critic", ag.target_critic)):
                for k_, v in flat_params(net).items():
                    out[f"s{si}_{tag}.{k_}"] = v
    np.savez_compressed(os.path.join(HERE, f"ddpg_{name}.npz"), **out)
    print(f"ddpg_{name}: {len(steps)} steps; last info {out[f's{len(steps) - 1}_info']}")


def td3_case(agent_mod, utils_mod, name, *, D, A, H, L, B, steps, seed, gamma=0.98, tau=0.05,
             grad_clip=1.0, lr=1e-3, policy_noise=0.2, noise_clamp=0.5, ac_update_freq=2, store_batches=True):
    """Unmodified reference TD3Agent.update on explicit batches; every torch.randn_like draw of
    the target-policy smoothing (src/agent.py:175) is recorded."""
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import ddpg as O
    torch.set_num_threads(1)
    cfg = utils_mod.BaseAgentConfig(
        hidden_dim=H, layer_count=L, actor_lr=lr, actor_lr_min=lr, ac_scheduler_steps=1, critic_lr=lr,
        critic_lr_min=lr, cr_scheduler_steps=1, buffer_type="HER", max_len=1000, alpha=1.0, batch_size=B,
        gamma=gamma, ac_update_freq=ac_update_freq, noise_std=0.2, noise_clamp=noise_clamp,
        policy_noise=policy_noise, grad_clip=grad_clip, beta=1.0, beta_end=1, k_future=4, max_eps_len=50,
        tau=tau)
    ag = agent_mod.TD3Agent(obs_dim=D, ac_dim=A, config=cfg, weights=None, nenvs=1, gradient_step=40)
    ag.device = "cpu"
    rng = np.random.default_rng(seed)
    nets0 = {"actor": O.init_mlp(rng, D, H, A, L), "critic_1": O.init_mlp(rng, D + A, H, 1, L),
             "critic_2": O.init_mlp(rng, D + A, H, 1, L)}
    with torch.no_grad():
        for tag, pref in (("actor", "base_net"), ("critic_1", "net"), ("critic_2", "net")):
            sd = getattr(ag, tag).state_dict()
            for i, (w, b) in enumerate(nets0[tag]):
                sd[f"{pref}.{2 * i}.weight"].copy_(torch.from_numpy(w))
                sd[f"{pref}.{2 * i}.bias"].copy_(torch.from_numpy(b))
    ag.update_target_network()
    out = {"meta": np.array([D, A, H, L, B, seed, ac_update_freq], np.int64),
           "hp": np.array([gamma, tau, grad_clip, lr, policy_noise, noise_clamp], np.float64),
           "steps": np.array(steps, np.int64)}
    torch.manual_seed(seed)
    real_randn_like = torch.randn_like
    for si, step in enumerate(steps):
        s = rng.standard_normal((B, D)).astype(np.float32)
        ns = (s + 0.1 * rng.standard_normal((B, D))).astype(np.float32)
        a = rng.uniform(-1, 1, (B, A)).astype(np.float32)
        r = -(rng.random((B, 1)) > 0.3).astype(np.float32)
        d = (rng.random((B, 1)) < 0.1).astype(np.float32)
        batch = tuple(torch.from_numpy(x) for x in (s, a, r, ns, d))
        ag.buffer.sample = lambda bs, _b=batch: _b
        drawn = []

        def rec(t, *a_, **k_):
            v = real_randn_like(t, *a_, **k_)
            drawn.append(v.numpy().copy())
            return v

        agent_mod.torch.randn_like = rec
        try:
            info = ag.update(step)
        finally:
            agent_mod.torch.randn_like = real_randn_like
        assert len(drawn) == 1
        if store_batches:
            for key, val in zip(("s", "a", "r", "ns", "d"), (s, a, r, ns, d)):
                out[f"s{si}_batch_{key}"] = val
        else:
            out[f"s{si}_batch_sum"] = np.array([x.astype(np.float64).sum() for x in (s, a, r, ns, d)])
        out[f"s{si}_noise"] = drawn[0]
        out[f"s{si}_info"] = np.array([float(x) for x in info], np.float64)
        if si == len(steps) - 1:
            for tag in ("actor", "critic_1", "critic_2", "target_actor", "target_critic_1", "target_critic_2"):
                for k_, v in flat_params(getattr(ag, tag)).items():
                    out[f"s{si}_{tag}.{k_}"] = v
    np.savez_compressed(os.path.join(HERE, f"td3_{name}.npz"), **out)
    print(f"td3_{name}: {len(steps)} steps; last info {out[f's{len(steps) - 1}_info']}")


def sac_case(agent_mod, utils_mod, algo, name, *, D, A, H, L, B, steps, seed, gamma=0.98, tau=0.05,
             grad_clip=1.0, lr=1e-3, alpha_lr=3e-4, alpha_min_steps=2, gradient_step=2, ac_update_freq=1):
    """Unmodified reference SACAgent / TQCAgent.update on explicit batches; every standard-normal
    draw behind ``Normal.rsample`` (src/model.py:134) is recorded: one for the next-state sample in
    critic_update, one for the state sample in actor_update."""
    import torch
    import torch.distributions.normal as tdn
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import ddpg as O
    from oracle import sac as OS
    torch.set_num_threads(1)
    cfg = utils_mod.SACAgentConfig(
        hidden_dim=H, layer_count=L, actor_lr=lr, actor_lr_min=lr, ac_scheduler_steps=1, critic_lr=lr,
        critic_lr_min=lr, cr_scheduler_steps=1, buffer_type="HER", max_len=1000, alpha=1.0, batch_size=B,
        gamma=gamma, ac_update_freq=ac_update_freq, noise_std=0.2, noise_clamp=0.5, policy_noise=0.2,
        grad_clip=grad_clip, beta=1.0, beta_end=1, k_future=4, max_eps_len=50, tau=tau, alpha_lr=alpha_lr,
        alpha_min_steps=alpha_min_steps)
    cls = agent_mod.SACAgent if algo == "sac" else agent_mod.TQCAgent
    ag = cls(obs_dim=D, ac_dim=A, config=cfg, weights=None, nenvs=1, gradient_step=gradient_step)
    ag.device = "cpu"
    rng = np.random.default_rng(seed)
    n = 2 if algo == "sac" else 5
    actor0, _ = OS.init_sac_actor(rng, D, H, A, L, head_scale=0.1, log_std_bias=-1.0)
    critics0 = [O.init_mlp(rng, D + A, H, 1, L) for _ in range(n)]
    critics = [ag.critic_1, ag.critic_2] if algo == "sac" else list(ag.critics)
    with torch.no_grad():
        sd = ag.actor.state_dict()
        for l in range(L):
            sd[f"base_net.{3 * l}.weight"].copy_(torch.from_numpy(actor0[2 * l][0]))
            sd[f"base_net.{3 * l}.bias"].copy_(torch.from_numpy(actor0[2 * l][1]))
        sd["mean_head.weight"].copy_(torch.from_numpy(actor0[2 * L][0]))
        sd["mean_head.bias"].copy_(torch.from_numpy(actor0[2 * L][1]))
        sd["log_std_head.weight"].copy_(torch.from_numpy(actor0[2 * L + 1][0]))
        sd["log_std_head.bias"].copy_(torch.from_numpy(actor0[2 * L + 1][1]))
        for c, p0 in zip(critics, critics0):
            sd = c.state_dict()
            for i, (w, b) in enumerate(p0):
                sd[f"net.{2 * i}.weight"].copy_(torch.from_numpy(w))
                sd[f"net.{2 * i}.bias"].copy_(torch.from_numpy(b))
    ag.update_target_network()
    out = {"meta": np.array([D, A, H, L, B, seed, ac_update_freq, gradient_step, alpha_min_steps], np.int64),
           "hp": np.array([gamma, tau, grad_clip, lr, alpha_lr], np.float64),
           "steps": np.array(steps, np.int64)}
    torch.manual_seed(seed)
    real = tdn._standard_normal
    for si, step in enumerate(steps):
        s = rng.standard_normal((B, D)).astype(np.float32)
        ns = (s + 0.1 * rng.standard_normal((B, D))).astype(np.float32)
        a = rng.uniform(-1, 1, (B, A)).astype(np.float32)
        r = -(rng.random((B, 1)) > 0.3).astype(np.float32)
        d = (rng.random((B, 1)) < 0.1).astype(np.float32)
        batch = tuple(torch.from_numpy(x) for x in (s, a, r, ns, d))
        ag.buffer.sample = lambda bs, _b=batch: _b
        drawn = []

        def rec(*a_, **k_):
            v = real(*a_, **k_)
            drawn.append(v.numpy().copy())
            return v

        tdn._standard_normal = rec
        try:
            info = ag.update(step)
        finally:
            tdn._standard_normal = real
        assert len(drawn) == (2 if step % ac_update_freq == 0 else 1), len(drawn)
        for key, val in zip(("s", "a", "r", "ns", "d"), (s, a, r, ns, d)):
            out[f"s{si}_batch_{key}"] = val
        out[f"s{si}_eps_next"] = drawn[0]
        out[f"s{si}_eps_cur"] = drawn[1] if len(drawn) > 1 else np.zeros_like(drawn[0])
        out[f"s{si}_info"] = np.array([float(x) for x in info], np.float64)
        out[f"s{si}_log_alpha"] = ag.log_alpha.detach().numpy().copy()
        if si == len(steps) - 1:
            tags = [("actor", ag.actor)]
            if algo == "sac":
                tags += [("critic_1", ag.critic_1), ("critic_2", ag.critic_2),
                         ("target_critic_1", ag.target_critic_1), ("target_critic_2", ag.target_critic_2)]
            else:
                keep = (0, len(ag.critics) - 1)      # first and last critic keep the fixture small
                tags += [(f"critic_{i}", c) for i, c in enumerate(ag.critics) if i in keep]
                tags += [(f"target_critic_{i}", c) for i, c in enumerate(ag.target_critics) if i in keep]
            for tag, net in tags:
                for k_, v in flat_params(net).items():
                    out[f"s{si}_{tag}.{k_}"] = v
    # deterministic / eval-mode action (select_action(eval_action=True), running statistics)
    x = rng.standard_normal((7, D)).astype(np.float32)
    out["eval_x"] = x
    out["eval_act"] = ag.select_action(x, eval_action=True)
    np.savez_compressed(os.path.join(HERE, f"{algo}_{name}.npz"), **out)
    print(f"{algo}_{name}: {len(steps)} steps; last info {out[f's{len(steps) - 1}_info']}")


def per_case(agent_mod, utils_mod, algo, name, *, D, A, H, L, B, max_len, n0, push_per_step, steps, seed,
             per_alpha=0.6, beta=0.4, beta_end=4, gamma=0.98, tau=0.05, grad_clip=1.0, lr=1e-3, ac_update_freq=1,
             gradient_step=2, store_final=True):
    """The prioritised-replay branch of update() (src/agent.py:1380-1387 and its twins), end to end through the
    UNMODIFIED reference: PERBuffer.push / sample / update_priorities (src/buffer.py:38-89) behind
    DDPG / TD3Agent / SACAgent / TQCAgent.update.  Recorded per step: the priorities before the draw, the
    uniforms np.random.choice consumed, the drawn positions, importance weights and batch, every torch normal
    draw, the returned tuple (per-sample TD errors separately) and the priorities afterwards."""
    import torch
    import torch.distributions.normal as tdn
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle import ddpg as O
    from oracle import sac as OS
    torch.set_num_threads(1)
    common = dict(hidden_dim=H, layer_count=L, actor_lr=lr, actor_lr_min=lr, ac_scheduler_steps=1, critic_lr=lr,
                  critic_lr_min=lr, cr_scheduler_steps=1, buffer_type="PER", max_len=max_len, alpha=per_alpha,
                  batch_size=B, gamma=gamma, ac_update_freq=ac_update_freq, noise_std=0.2, noise_clamp=0.5,
                  policy_noise=0.2, grad_clip=grad_clip, beta=beta, beta_end=beta_end, k_future=4, max_eps_len=50,
                  tau=tau)
    rng = np.random.default_rng(seed)
    if algo in ("ddpg", "td3"):
        cfg = utils_mod.BaseAgentConfig(**common)
        cls = agent_mod.DDPG if algo == "ddpg" else agent_mod.TD3Agent
        ag = cls(obs_dim=D, ac_dim=A, config=cfg, weights=None, nenvs=1, gradient_step=40)
        nets = {"actor": ("base_net", O.init_mlp(rng, D, H, A, L))}
        for tag in (("critic",) if algo == "ddpg" else ("critic_1", "critic_2")):
            nets[tag] = ("net", O.init_mlp(rng, D + A, H, 1, L))
        with torch.no_grad():
            for tag, (pref, params) in nets.items():
                sd = getattr(ag, tag).state_dict()
                for i, (w, b) in enumerate(params):
                    sd[f"{pref}.{2 * i}.weight"].copy_(torch.from_numpy(w))
                    sd[f"{pref}.{2 * i}.bias"].copy_(torch.from_numpy(b))
        final_tags = list(nets) + ["target_" + t for t in nets]
    else:
        cfg = utils_mod.SACAgentConfig(**common, alpha_lr=1e-2, alpha_min_steps=0)
        cls = agent_mod.SACAgent if algo == "sac" else agent_mod.TQCAgent
        ag = cls(obs_dim=D, ac_dim=A, config=cfg, weights=None, nenvs=1, gradient_step=gradient_step)
        n = 2 if algo == "sac" else 5
        actor0, _ = OS.init_sac_actor(rng, D, H, A, L, head_scale=0.1, log_std_bias=-1.0)
        critics0 = [O.init_mlp(rng, D + A, H, 1, L) for _ in range(n)]
        critics = [ag.critic_1, ag.critic_2] if algo == "sac" else list(ag.critics)
        with torch.no_grad():
            sd = ag.actor.state_dict()
            for l in range(L):
                sd[f"base_net.{3 * l}.weight"].copy_(torch.from_numpy(actor0[2 * l][0]))
                sd[f"base_net.{3 * l}.bias"].copy_(torch.from_numpy(actor0[2 * l][1]))
            sd["mean_head.weight"].copy_(torch.from_numpy(actor0[2 * L][0]))
            sd["mean_head.bias"].copy_(torch.from_numpy(actor0[2 * L][1]))
            sd["log_std_head.weight"].copy_(torch.from_numpy(actor0[2 * L + 1][0]))
            sd["log_std_head.bias"].copy_(torch.from_numpy(actor0[2 * L + 1][1]))
            for c, p0 in zip(critics, critics0):
                sd = c.state_dict()
                for i, (w, b) in enumerate(p0):
                    sd[f"net.{2 * i}.weight"].copy_(torch.from_numpy(w))
                    sd[f"net.{2 * i}.bias"].copy_(torch.from_numpy(b))
        final_tags = None
    ag.device = "cpu"
    ag.buffer.device = "cpu"
    ag.update_target_network()
    assert type(ag.buffer).__name__ == "PERBuffer"

    pushed = {k: [] for k in ("s", "a", "r", "ns", "d")}

    def push(count):
        for _ in range(count):
            s = rng.standard_normal(D).astype(np.float32)
            ns = (s + 0.1 * rng.standard_normal(D)).astype(np.float32)
            a = rng.uniform(-1, 1, A).astype(np.float32)
            r = np.float64(-float(rng.random() > 0.3))
            d = np.bool_(rng.random() < 0.1)
            ag.push(torch.from_numpy(s), a, r, torch.from_numpy(ns), d)       # the types src/env.py:226 hands over
            for k, v in zip(("s", "a", "r", "ns", "d"), (s, a, np.float32(r), ns, np.float32(d))):
                pushed[k].append(v)

    out = {"meta": np.array([D, A, H, L, B, seed, ac_update_freq, gradient_step, max_len, n0, push_per_step, beta_end],
                            np.int64),
           "hp": np.array([gamma, tau, grad_clip, lr, per_alpha, beta], np.float64),
           "steps": np.array(steps, np.int64)}
    np.random.seed(seed)
    torch.manual_seed(seed)
    real_choice, real_normal, real_randn_like = np.random.choice, tdn._standard_normal, torch.randn_like
    real_sample = ag.buffer.sample
    push(n0)
    for si, step in enumerate(steps):
        if si:
            push(push_per_step)
        out[f"s{si}_prio_before"] = np.array(ag.buffer.priorities, dtype=np.float32)
        out[f"s{si}_beta"] = np.float64(ag.beta)
        rec = {"normal": []}

        def choice(N, size, p=None):
            st = np.random.get_state()
            idx = real_choice(N, size, p=p)
            after = np.random.get_state()
            np.random.set_state(st)
            rec["u"] = np.random.random_sample(size)
            assert all(np.array_equal(x, y) for x, y in zip(np.random.get_state(), after)), "choice drew more"
            rec["P"] = np.array(p, copy=True)
            return idx

        def sample(bs, b):
            res = real_sample(bs, b)
            rec["batch"] = [t.numpy().copy() for t in res[:5]]
            rec["w"], rec["idx"] = res[5].numpy().copy(), np.asarray(res[6]).copy()
            return res

        def normal(*a_, **k_):
            v = real_normal(*a_, **k_)
            rec["normal"].append(v.numpy().copy())
            return v

        def randn_like(t, *a_, **k_):
            v = real_randn_like(t, *a_, **k_)
            rec["normal"].append(v.numpy().copy())
            return v

        np.random.choice, ag.buffer.sample = choice, sample
        tdn._standard_normal, agent_mod.torch.randn_like = normal, randn_like
        try:
            info = ag.update(step)
        finally:
            np.random.choice, tdn._standard_normal, agent_mod.torch.randn_like = real_choice, real_normal, real_randn_like
            ag.buffer.sample = real_sample
        td_pos = {"ddpg": 2 if len(info) == 6 else 1, "td3": 3 if len(info) == 8 else 2}.get(algo, 3 if len(info) == 9 else 2)
        td = np.asarray(info[td_pos], np.float32).reshape(B, 1)
        flat = [float(np.mean(x)) if i == td_pos else float(x) for i, x in enumerate(info)]
        out[f"s{si}_info"] = np.array(flat, np.float64)
        out[f"s{si}_td"] = td
        out[f"s{si}_u"], out[f"s{si}_P"] = rec["u"], rec["P"].astype(np.float32)
        out[f"s{si}_idx"], out[f"s{si}_w"] = rec["idx"].astype(np.int64), rec["w"].astype(np.float32)
        for key, val in zip(("s", "a", "r", "ns", "d"), rec["batch"]):
            out[f"s{si}_batch_{key}"] = val
        for j, v in enumerate(rec["normal"]):
            out[f"s{si}_normal{j}"] = v
        out[f"s{si}_n_normal"] = np.int64(len(rec["normal"]))
        out[f"s{si}_prio_after"] = np.array(ag.buffer.priorities, dtype=np.float32)
        if algo in ("sac", "tqc"):
            out[f"s{si}_log_alpha"] = ag.log_alpha.detach().numpy().copy()
    for k in pushed:
        out["push_" + k] = np.stack(pushed[k]).astype(np.float32)
    si = len(steps) - 1
    if not store_final:
        pass
    elif final_tags is not None:
        for tag in final_tags:
            for k_, v in flat_params(getattr(ag, tag)).items():
                out[f"s{si}_{tag}.{k_}"] = v
    else:
        tags = [("actor", ag.actor)]
        if algo == "sac":
            tags += [("critic_1", ag.critic_1), ("critic_2", ag.critic_2), ("target_critic_1", ag.target_critic_1),
                     ("target_critic_2", ag.target_critic_2)]
        else:
            keep = (0, len(ag.critics) - 1)
            tags += [(f"critic_{i}", c) for i, c in enumerate(ag.critics) if i in keep]
            tags += [(f"target_critic_{i}", c) for i, c in enumerate(ag.target_critics) if i in keep]
        for tag, net in tags:
            for k_, v in flat_params(net).items():
                out[f"s{si}_{tag}.{k_}"] = v
    np.savez_compressed(os.path.join(HERE, f"per_{algo}_{name}.npz"), **out)
    print(f"per_{algo}_{name}: {len(steps)} steps; len {len(ag.buffer)}; last info {out[f's{si}_info']}")


def per_cases(agent_mod, utils_mod):
    per_case(agent_mod, utils_mod, "ddpg", "reach_h64", D=10, A=3, H=64, L=3, B=64, max_len=500, n0=300,
             push_per_step=90, steps=[38, 39, 40, 41, 42], seed=41, grad_clip=10.0)
    per_case(agent_mod, utils_mod, "ddpg", "push_h256", D=21, A=3, H=256, L=3, B=256, max_len=3000, n0=3000,
             push_per_step=7, steps=[39, 40, 41], seed=42, beta_end=41, ac_update_freq=2, store_final=False)
    per_case(agent_mod, utils_mod, "td3", "push_h64", D=22, A=3, H=64, L=3, B=128, max_len=700, n0=400,
             push_per_step=200, steps=[1, 2, 3, 4], seed=43, ac_update_freq=2)
    per_case(agent_mod, utils_mod, "sac", "push_h64", D=22, A=3, H=64, L=3, B=128, max_len=700, n0=700,
             push_per_step=50, steps=[1, 2, 3, 4], seed=44)
    per_case(agent_mod, utils_mod, "tqc", "slide_h64", D=22, A=3, H=64, L=3, B=128, max_len=600, n0=200,
             push_per_step=150, steps=[1, 2, 3], seed=45)


def checkpoint_case(model_mod):
    """Load-compat + forward-differential fixture from the shipped Reach checkpoint
    (resources/DDPG/reach/{actor,critic}.pth: H=64, D=10, A=3)."""
    import torch
    sd_a = torch.load(os.path.join(REF, "resources/DDPG/reach/actor.pth"), map_location="cpu")
    sd_c = torch.load(os.path.join(REF, "resources/DDPG/reach/critic.pth"), map_location="cpu")
    D = sd_a["base_net.0.weight"].shape[1]
    H = sd_a["base_net.0.weight"].shape[0]
    A = sd_a["base_net.6.weight"].shape[0]
    actor = model_mod.Actor(D, H, A, 3)
    critic = model_mod.Critic(D + A, H, 3)
    actor.load_state_dict(sd_a)
    critic.load_state_dict(sd_c)
    rng = np.random.default_rng(3)
    x = rng.standard_normal((37, D)).astype(np.float32)
    with torch.no_grad():
        act = actor(torch.from_numpy(x))
        q = critic(torch.cat([torch.from_numpy(x), act], -1))
    out = {"x": x, "act": act.numpy(), "q": q.numpy()}
    for k, v in sd_a.items():
        out["actor." + k] = v.numpy()
    for k, v in sd_c.items():
        out["critic." + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "checkpoint_reach.npz"), **out)
    print("checkpoint_reach: D,H,A =", D, H, A)


def sac_cases(agent_mod, utils_mod):
    sac_case(agent_mod, utils_mod, "sac", "push_h64", D=22, A=3, H=64, L=3, B=128, steps=[1, 2, 3, 4, 5],
             seed=31)
    sac_case(agent_mod, utils_mod, "sac", "pickplace_h256", D=23, A=4, H=256, L=2, B=200, steps=[4, 5, 6],
             seed=32, grad_clip=0.1, tau=0.005, ac_update_freq=2, gradient_step=3, alpha_min_steps=0)
    sac_case(agent_mod, utils_mod, "tqc", "slide_h64", D=22, A=3, H=64, L=3, B=128, steps=[1, 2, 3, 4, 5],
             seed=33)
    sac_case(agent_mod, utils_mod, "tqc", "push_h256", D=22, A=3, H=256, L=3, B=256, steps=[9, 10, 11],
             seed=34, grad_clip=0.5, tau=0.005, alpha_min_steps=0, alpha_lr=1e-2)


def large_batch_cases(agent_mod, utils_mod):
    """Large-batch updates on the PickAndPlace shape (time-feature obs 20 + goal 3 = 23, act 4, H 256 x 3, k 8:
    src/config/DDPG/config_ddpg_pickplace.yaml:14-45) -- the batch sizes the tensor-core engine serves
    (>= 2048).  Batches are regenerated from the seed by the tests; weights are stored for the last step."""
    ddpg_case(agent_mod, utils_mod, "pickplace_B2048", D=23, A=4, H=256, L=3, B=2048,
              steps=[39, 40, 41], seed=31, store_weights=False, store_batches=False)
    ddpg_case(agent_mod, utils_mod, "pickplace_B8192", D=23, A=4, H=256, L=3, B=8192,
              steps=[40, 41, 42], seed=32, store_weights=False, store_batches=False)
    td3_case(agent_mod, utils_mod, "pickplace_B4096", D=23, A=4, H=256, L=3, B=4096, steps=[1, 2, 3], seed=33,
             ac_update_freq=2, store_batches=False)


def main():
    agent_mod, buffer_mod, model_mod, utils_mod = import_reference()
    if len(sys.argv) > 1 and sys.argv[1] == "sac":
        return sac_cases(agent_mod, utils_mod)
    if len(sys.argv) > 1 and sys.argv[1] == "per":
        return per_cases(agent_mod, utils_mod)
    if len(sys.argv) > 1 and sys.argv[1] == "large":
        return large_batch_cases(agent_mod, utils_mod)
    td3_case(agent_mod, utils_mod, "push_h64", D=22, A=3, H=64, L=3, B=128, steps=[1, 2, 3, 4, 5], seed=21)
    td3_case(agent_mod, utils_mod, "pickplace_h256", D=23, A=4, H=256, L=2, B=200, steps=[7, 8, 9], seed=22,
             grad_clip=0.1, ac_update_freq=1, tau=0.005, policy_noise=0.3, noise_clamp=0.25)
    if len(sys.argv) > 1 and sys.argv[1] == "td3":
        return
    reward_case()
    her_case(buffer_mod, "reach_small", O=7, G=3, A=3, k=4,
             lens=[50, 50, 7, 1, 2, 50, 13, 50], max_mem_len=100000, nenvs=2, seed=0,
             batches=[32, 256, 1])
    her_case(buffer_mod, "push_evict", O=19, G=3, A=3, k=4,
             lens=[50, 50, 50, 3, 50, 50, 21, 50, 50], max_mem_len=777, nenvs=3, seed=1,
             batches=[64, 512])
    her_case(buffer_mod, "pickplace_k8", O=20, G=3, A=4, k=8,
             lens=[50, 50, 50, 50, 9, 50], max_mem_len=100000, nenvs=4, seed=2,
             batches=[128, 1000])
    her_case(buffer_mod, "k0", O=6, G=3, A=3, k=0,
             lens=[50, 5, 50], max_mem_len=1000, nenvs=1, seed=3, batches=[50])
    normalizer_case(utils_mod)
    ddpg_case(agent_mod, utils_mod, "reach_h64", D=10, A=3, H=64, L=3, B=64,
              steps=[38, 39, 40, 41, 42], seed=5)
    ddpg_case(agent_mod, utils_mod, "push_h256", D=21, A=3, H=256, L=3, B=256,
              steps=[39, 40, 41], seed=6, store_weights=False)
    ddpg_case(agent_mod, utils_mod, "pickplace_l2_cosine", D=23, A=4, H=128, L=2, B=96,
              steps=[1, 2, 3, 4, 5, 6, 7, 8], seed=8, actor_lr=1e-3, actor_lr_min=1e-4, ac_T=3,
              critic_lr=2e-3, critic_lr_min=5e-4, cr_T=2, ac_update_freq=2, gamma=0.95,
              grad_clip=0.05, tau=0.005, store_weights=False)
    checkpoint_case(model_mod)
    sac_cases(agent_mod, utils_mod)
    per_cases(agent_mod, utils_mod)
    large_batch_cases(agent_mod, utils_mod)


if __name__ == "__main__":
    main()
