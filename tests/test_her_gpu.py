"""GPU parity: device-resident HER store + lazy relabelling sampler (csrc/her.cu) through the
C ABI versus the oracle / the fixtures dumped from the unmodified reference.

Bar: relabelled goals, rewards (incl. the -0.0 sign), dones and sample indices BIT-EXACT.
"""
import random

import numpy as np
import pytest

from oracle import her as OH
from oracle.index_stream import feistel_positions
from tests.helpers import HER_CASES, bits, her_episodes, load

pytestmark = pytest.mark.gpu

HER_SEEDS = {"reach_small": 0, "push_evict": 1, "pickplace_k8": 2, "k0": 3}   # make_golden.py::main

FIELDS = ("s", "a", "r", "ns", "d")


def make_buffer(g, **kw):
    from gcrl_b200 import HERBuffer
    O, G, A, k, max_len, nenvs = (int(x) for x in g["meta"])
    return HERBuffer(max_len, 50, nenvs, k_future=k, **kw), k


def fill(buf, eps):
    for ep in eps:
        buf.push_episode(ep["s"], ep["a"], ep["ns"], ep["r"], ep["d"], ep["ag"], ep["fut"])


def assert_bits(got, ref, what):
    got = np.asarray(got, np.float32).reshape(np.asarray(ref).shape)
    assert np.array_equal(bits(got), bits(ref)), what


@pytest.mark.parametrize("case", HER_CASES)
def test_full_dump_equals_reference_deque(case):
    """Sampling positions 0..len-1 reproduces the reference deque entry for entry."""
    g = load("her_" + case)
    buf, _ = make_buffer(g)
    fill(buf, her_episodes(g))
    n = int(g["len"])
    assert len(buf) == n
    out = buf.sample_host(n, indices=np.arange(n))
    for got, key in zip(out, ("dump_s", "dump_a", "dump_r", "dump_ns", "dump_d")):
        assert_bits(got, g[key], (case, key))


@pytest.mark.parametrize("case", HER_CASES)
def test_sample_matches_reference_batches(case):
    """The reference's own sample() outputs on its recorded random.sample index stream."""
    g = load("her_" + case)
    buf, _ = make_buffer(g)
    fill(buf, her_episodes(g))
    for bi in range(int(g["n_batches"])):
        idx = g[f"b{bi}_idx"]
        out = buf.sample(len(idx), indices=idx)               # device tensors
        for got, key in zip(out, FIELDS):
            assert got.is_cuda and got.dtype.is_floating_point
            assert tuple(got.shape) == g[f"b{bi}_{key}"].shape
            assert_bits(got.cpu().numpy(), g[f"b{bi}_{key}"], (case, bi, key))


@pytest.mark.parametrize("case", ["reach_small", "push_evict"])
def test_push_api_consumes_reference_mt_stream(case):
    """push() per transition, then sample(): seeded like the fixture generator (random.seed(1898 + seed),
    tests/golden/make_golden.py::her_case) the product must consume the interpreter's Mersenne-Twister exactly
    as the reference does -- every random.randint of apply_her (src/buffer.py:153) and every random.sample of
    sample() (:124), now drawn by the C mirror of CPython's generator -- so the deque dump AND the sampled
    batches are reproduced bit for bit without being told a single index."""
    g = load("her_" + case)
    eps = her_episodes(g)
    buf, k = make_buffer(g)
    random.seed(1898 + HER_SEEDS[case])
    for ep in eps:                                             # commit order == golden order
        for t in range(ep["s"].shape[0]):
            buf.push(0, ep["s"][t], ep["a"][t], ep["ns"][t], ep["r"][t], bool(ep["d"][t]),
                     ep["dg"][t], ep["ag"][t])
    n = int(g["len"])
    out = buf.sample_host(n, indices=np.arange(n))
    for got, key in zip(out, ("dump_s", "dump_a", "dump_r", "dump_ns", "dump_d")):
        assert_bits(got, g[key], (case, key))
    for bi in range(int(g["n_batches"])):                      # the reference's sample() calls, in order
        B = g[f"b{bi}_idx"].shape[0]
        *batch, used = buf.sample_host(B, return_indices=True)
        assert np.array_equal(used, g[f"b{bi}_idx"]), (case, bi)
        for got, key in zip(batch, ("s", "a", "r", "ns", "d")):
            assert_bits(got, g[f"b{bi}_{key}"], (case, bi, key))


def test_reward_known_answers_through_the_kernel():
    """Threshold edge cases (exactly 0.05f, nextafter either side) and 4096 random pairs:
    T=2 episodes with ag = (a, b), k=1, fut=1 -> entry 1 carries reward(a, b)."""
    from gcrl_b200 import HERBuffer
    g = load("reward_kat")
    a, b, r = g["a"], g["b"], g["r"]
    n = a.shape[0]
    buf = HERBuffer(10 * n, 50, 1, k_future=1)
    rng = np.random.default_rng(0)
    for i in range(n):
        s = rng.standard_normal((2, 6)).astype(np.float32)
        buf.push_episode(s, np.zeros((2, 3), np.float32), s, np.full(2, -1.0, np.float32),
                         np.zeros(2, np.float32), np.stack([a[i], b[i]]),
                         np.array([[1], [0]], np.uint8))
    assert len(buf) == 3 * n
    out = buf.sample_host(n, indices=np.arange(n) * 3 + 1)
    assert_bits(out[2][:, 0], r, "reward")
    assert_bits(out[0][:, -3:], b, "relabelled goal")
    assert_bits(out[3][:, -3:], b, "relabelled next goal")
    assert set(np.unique(bits(out[2]))) == {0x80000000, 0xBF800000}
    assert not out[4].any()


def test_underfilled_and_bad_arguments():
    from gcrl_b200 import HERBuffer
    buf = HERBuffer(1000, 50, 1)
    with pytest.raises(AssertionError):
        buf.sample(1)
    g = load("her_k0")
    buf, _ = make_buffer(g)
    fill(buf, her_episodes(g))
    with pytest.raises(AssertionError):
        buf.sample(len(buf) + 1)
    with pytest.raises(ValueError):
        buf.sample(4, indices=[0, 1, 2, len(buf)])
    g2 = load("her_reach_small")
    b2, _ = make_buffer(g2)
    e2 = her_episodes(g2)[0]
    bad = e2["fut"].copy()
    bad[3, 0] = 3                                              # future index must be in [t+1, T-1]
    with pytest.raises(ValueError):
        b2.push_episode(e2["s"], e2["a"], e2["ns"], e2["r"], e2["d"], e2["ag"], bad)
    assert len(b2) == 0                                        # a rejected episode leaves no trace


def test_device_index_stream_matches_restatement_and_is_distinct():
    g = load("her_pickplace_k8")
    buf, _ = make_buffer(g, index_source="device", seed=1898)
    fill(buf, her_episodes(g))
    n = len(buf)
    dump = {k: g["dump_" + k] for k in FIELDS}
    for epoch in range(3):
        *out, used = buf.sample_host(512, return_indices=True)
        want = feistel_positions(np.arange(512), n, 1898, epoch)
        assert np.array_equal(used, want), epoch
        assert len(set(used.tolist())) == 512                  # without replacement
        for got, key in zip(out, FIELDS):
            assert_bits(got, dump[key][used], (epoch, key))


def synth(rng, E, T, O, G, A):
    obs = rng.standard_normal((E, T + 1, O)).astype(np.float32)
    ag = np.empty((E, T + 1, G), np.float32)
    ag[:, 0] = rng.uniform(-0.15, 0.15, (E, G))
    steps = (rng.normal(0, 0.02, (E, T, G)) * (rng.random((E, T, 1)) > 0.3)).astype(np.float32)
    for t in range(T):
        ag[:, t + 1] = ag[:, t] + steps[:, t]
    dg = rng.uniform(-0.15, 0.15, (E, 1, G)).astype(np.float32)
    act = rng.uniform(-1, 1, (E, T, A)).astype(np.float32)
    s = np.concatenate([obs[:, :-1], np.repeat(dg, T, 1)], -1)
    ns = np.concatenate([obs[:, 1:], np.repeat(dg, T, 1)], -1)
    r = OH.compute_reward(ag[:, 1:], np.repeat(dg, T, 1))
    return s, act, ns, r, np.zeros((E, T), np.float32), ag[:, 1:]


@pytest.mark.parametrize("max_len,E", [(40_000, 300), (1_000_000, 700)])
def test_midsize_against_vectorised_oracle_with_eviction(max_len, E):
    """Hundreds of episodes (ragged lengths), per-entry FIFO eviction, both index sources."""
    from gcrl_b200 import HERBuffer
    rng = np.random.default_rng(42)
    O, G, A, k = 19, 3, 3, 4
    buf = HERBuffer(max_len, 50, 1, k_future=k)
    parts = []
    for e in range(E):
        T = 50 if rng.random() < 0.8 else int(rng.integers(1, 50))
        s, a, ns, r, d, ag = (x[0] for x in synth(rng, 1, T, O, G, A))
        if T < 50:
            d = d.copy()
            d[-1] = 1.0
        fut = np.zeros((T, k), np.uint8)
        for t in range(T - 1):
            fut[t] = rng.integers(t + 1, T, k)
        buf.push_episode(s, a, ns, r, d, ag, fut)
        parts.append(OH.materialise_episode(s, a, ns, r, d, ag, fut, k))
    full = [np.concatenate([p[i] for p in parts]) for i in range(5)]
    n = min(full[0].shape[0], max_len)
    assert len(buf) == n
    live = [x[-n:] for x in full]
    idx = rng.permutation(n)[:min(n, 30_000)]
    out = buf.sample_host(len(idx), indices=idx)
    for got, ref, key in zip(out, live, FIELDS):
        assert_bits(got, ref[idx], key)


def test_full_size_properties_1m_buffer():
    """BASELINE size (1M-entry buffer, T=50, k=4, Push shape): size-independent properties.
    (i) len == maxlen after overfilling; (ii) every sampled row is either an untouched
    original or a relabel whose goal columns equal an achieved goal of the SAME episode at a
    later step and whose reward is the sparse rule recomputed on the host; (iii) device
    stream indices are distinct and in range; (iv) sampling is idempotent."""
    from gcrl_b200 import HERBuffer
    rng = np.random.default_rng(7)
    O, G, A, k, T = 19, 3, 3, 4, 50
    per = (T - 1) * (k + 1) + 1
    E = 1_000_000 // per + 40
    s, a, ns, r, d, ag = synth(rng, E, T, O, G, A)
    fut = np.zeros((E, T, k), np.uint8)
    for t in range(T - 1):
        fut[:, t] = rng.integers(t + 1, T, (E, k))
    buf = HERBuffer(1_000_000, 50, 1, k_future=k, index_source="device")
    for e in range(E):
        buf.push_episode(s[e], a[e], ns[e], r[e], d[e], ag[e], fut[e])
    assert len(buf) == 1_000_000
    B = 65536
    *out, used = buf.sample_host(B, return_indices=True)
    assert used.min() >= 0 and used.max() < 1_000_000 and len(np.unique(used)) == B
    ge = E * per - 1_000_000 + used                     # global entry ids
    ep, o = ge // per, ge % per
    t = np.minimum(o // (k + 1), T - 1)
    j = np.where(o < (T - 1) * (k + 1), o % (k + 1), 0)
    rel = j > 0
    f = fut[ep[rel], t[rel], j[rel] - 1].astype(np.int64)
    assert (f > t[rel]).all()
    S, Aa, R, NS, Dn = out
    assert_bits(Aa, a[ep, t], "actions")
    assert_bits(S[:, :O], s[ep, t, :O], "obs")
    assert_bits(NS[:, :O], ns[ep, t, :O], "next obs")
    assert_bits(S[~rel], s[ep[~rel], t[~rel]], "originals")
    assert_bits(R[~rel, 0], r[ep[~rel], t[~rel]], "original rewards")
    assert_bits(S[rel, O:], ag[ep[rel], f], "relabelled goal")
    assert_bits(NS[rel, O:], ag[ep[rel], f], "relabelled next goal")
    assert_bits(R[rel, 0], OH.compute_reward(ag[ep[rel], t[rel]], ag[ep[rel], f]), "relabel reward")
    assert not Dn[rel].any()
    again = buf.sample_host(B, indices=used)
    for x, y in zip(out, again):
        assert np.array_equal(bits(x), bits(y))


def test_buffer_and_agent_resume_from_checkpoint_sample_the_same_batches(tmp_path):
    """A partly evicted buffer + per-env staging + the interpreter's random state + the agent survive
    save_checkpoint / load_checkpoint: the resumed run draws the same positions, gathers the same bits and
    makes bit-identical updates, including the episodes committed after the resume."""
    import torch
    from gcrl_b200 import DDPG
    from tests.test_ddpg_gpu import make_config
    g = load("her_push_evict")
    eps = her_episodes(g)
    D, A = 22, 3
    cfg = make_config(hidden_dim=64, batch_size=64, max_len=777)        # 777 entries: the oldest episodes are evicted

    def feed(ag, ep, upto=None):
        T = ep["s"].shape[0] if upto is None else upto
        for t in range(T):
            ag.push_her(0, ep["s"][t], ep["a"][t], ep["ns"][t], ep["r"][t], bool(ep["d"][t]), ep["dg"][t], ep["ag"][t])
    torch.manual_seed(1)
    random.seed(77)
    a1 = DDPG(D, A, cfg, None, 1, 40)
    for ep in eps[:6]:
        feed(a1, ep)
    feed(a1, eps[6], upto=9)                                            # 9 transitions left in the staging deque
    for step in (1, 2, 3):
        a1.update(step)
    a1.save_checkpoint(str(tmp_path / "ck"))

    def rest(ag):
        out = []
        ep = eps[6]
        for t in range(9, ep["s"].shape[0]):
            ag.push_her(0, ep["s"][t], ep["a"][t], ep["ns"][t], ep["r"][t], bool(ep["d"][t]), ep["dg"][t], ep["ag"][t])
        feed(ag, eps[7])
        for step in (4, 5, 6):
            out.append([float(x) for x in ag.update(step)])
        n = len(ag.buffer)
        return out, n, [x.copy() for x in ag.buffer.sample_host(n, indices=np.arange(n))]
    ref = rest(a1)
    torch.manual_seed(2)
    random.seed(5)
    a2 = DDPG(D, A, cfg, None, 1, 40)
    a2.load_checkpoint(str(tmp_path / "ck"))
    got = rest(a2)
    assert ref[0] == got[0] and ref[1] == got[1]
    for x, y in zip(ref[2], got[2]):
        assert_bits(x, y, "resumed buffer dump")


@pytest.mark.parametrize("threshold", [0.05, 0.02, 0.11])
def test_threshold_and_injected_reward_are_honoured(threshold):
    """HERBuffer(threshold=...) reaches the kernel (relabelled rewards follow -(d > threshold), bit for bit), a
    compute_reward that IS that rule is accepted, anything else is rejected at the first commit."""
    from gcrl_b200 import HERBuffer
    rng = np.random.default_rng(11)
    O, G, A, k, T, E = 7, 3, 3, 4, 50, 6
    s, a, ns, r, d, ag = synth(rng, E, T, O, G, A)
    fut = np.zeros((E, T, k), np.uint8)
    for t in range(T - 1):
        fut[:, t] = rng.integers(t + 1, T, (E, k))
    thr32 = np.float32(threshold)

    def rule(x, y, info=None):                           # the oracle's arithmetic with this threshold
        diff = np.asarray(x, np.float32) - np.asarray(y, np.float32)
        sq = diff * diff
        acc = sq[..., 0]
        for c in range(1, sq.shape[-1]):
            acc = acc + sq[..., c]
        return -(np.sqrt(acc) > thr32).astype(np.float32)

    buf = HERBuffer(100_000, 50, 1, threshold=threshold, k_future=k)
    buf.compute_reward = rule                            # assigned before the first transition, like src/env.py:105
    for e in range(E):
        buf.push_episode(s[e], a[e], ns[e], r[e], d[e], ag[e], fut[e])
    per = (T - 1) * (k + 1) + 1
    idx = np.arange(E * per)
    S, Aa, R, NS, Dn = buf.sample_host(len(idx), indices=idx)
    ep, o = idx // per, idx % per
    t = np.minimum(o // (k + 1), T - 1)
    j = np.where(o < (T - 1) * (k + 1), o % (k + 1), 0)
    rel = j > 0
    f = fut[ep[rel], t[rel], j[rel] - 1].astype(np.int64)
    want = rule(ag[ep[rel], t[rel]], ag[ep[rel], f])
    assert_bits(R[rel, 0], want, "relabel reward at this threshold")
    assert 0 < np.count_nonzero(want) < want.size        # both outcomes occur, so the threshold matters
    bad = HERBuffer(100_000, 50, 1, threshold=threshold, k_future=k)
    bad.compute_reward = lambda x, y, info: -np.float32(np.linalg.norm(x - y))      # dense reward
    with pytest.raises(ValueError, match="compute_reward"):
        bad.push_episode(s[0], a[0], ns[0], r[0], d[0], ag[0], fut[0])
