"""oracle/per.py pinned three ways (CPU only):
  * its restatements of NumPy's float32 pairwise sum and of RandomState.choice against NumPy itself;
  * ReplayBufferOracle / PERBufferOracle against fixtures dumped from the unmodified reference classes behind the
    unmodified reference agents (tests/golden/make_golden.py::per_case), step by step ("teacher forced": every
    step starts from the reference's recorded priorities, so one step's float32 ``power`` ulp cannot leak into
    the next step's positions);
  * the ``weights`` branch of the DDPG / TD3 / SAC / TQC update oracles against the same fixtures.
Tolerances: positions, rows, P and the searched table bit-exact; importance weights and new priorities 2 ulp of
float32 (platform powf); losses / Q / gradient norms rel 2e-5 + 1e-6; weights tests.helpers.weights_close."""
import random

import numpy as np
import pytest

from oracle import per as OP
from tests.helpers import (PER_CASES, assert_sac_actor_close, bits, ddpg_params_from_golden, load, per_meta, per_oracle,
                           per_pushes, per_td_position, sac_params_from_golden, ulp_diff_f32, weights_close)


@pytest.mark.parametrize("n", list(range(1, 140)) + [255, 256, 257, 1000, 1031, 4097, 65537, 100003])
def test_pairwise_sum_is_numpys(n):
    rng = np.random.default_rng(n)
    a = (rng.random(n) ** 4 * 10).astype(np.float32)
    assert bits(OP.pairwise_sum_f32(a)) == bits(a.sum())


@pytest.mark.parametrize("n", [1, 7, 8, 128, 129, 300, 1000, 4097, 100003])
def test_pairwise_tree_table_reproduces_the_sum(n):
    """The (leaf, node) table the CUDA library builds for the device-side fold follows the same recursion."""
    rng = np.random.default_rng(n + 1)
    a = (rng.random(n) * 3).astype(np.float32)
    leaves, nodes = OP.pairwise_leaves(n)
    vals = [OP.pairwise_sum_f32(a[lo:lo + m]) if m >= 8 or n < 8 else None for lo, m in leaves]
    assert all(v is not None for v in vals)
    assert sum(m for _, m in leaves) == n and all(m <= 128 for _, m in leaves)
    for l, r in nodes:
        vals.append(np.float32(vals[l] + vals[r]))
    assert bits(np.float32(np.float32(0) + vals[-1])) == bits(a.sum())


@pytest.mark.parametrize("n,B,seed", [(5, 3, 0), (300, 64, 1), (3000, 256, 2), (70001, 512, 3)])
def test_choice_restatement_is_numpys(n, B, seed):
    rng = np.random.default_rng(seed)
    prio = (rng.random(n) ** 3 + 1e-4).astype(np.float32)
    P = OP.normalised_priorities(prio)
    ref_P = prio.copy()
    ref_P /= ref_P.sum()
    assert np.array_equal(bits(P), bits(ref_P))
    np.random.seed(seed)
    st = np.random.get_state()
    want = np.random.choice(n, B, p=ref_P)
    np.random.set_state(st)
    u = np.random.random_sample(B)
    assert np.array_equal(OP.choice_indices(P, u), want)
    assert np.array_equal(OP.choice_cdf(P), (lambda c: c / c[-1])(ref_P.astype(np.float64).cumsum()))


def test_uniform_replay_oracle_follows_random_sample():
    rng = np.random.default_rng(0)
    buf = OP.ReplayBufferOracle(50)
    rows = []
    for i in range(80):
        row = (rng.standard_normal(4).astype(np.float32), rng.standard_normal(2).astype(np.float32), -float(i % 2),
               rng.standard_normal(4).astype(np.float32), float(i % 3 == 0))
        rows.append(row)
        buf.push(*row)
    assert len(buf) == 50
    random.seed(5)
    s, a, r, ns, d = buf.sample(16)
    random.seed(5)
    want = random.sample(range(50), 16)
    assert list(buf.last_indices) == want
    live = rows[30:]
    assert np.array_equal(s, np.stack([live[i][0] for i in want]))
    assert np.array_equal(r[:, 0], np.array([live[i][2] for i in want], np.float32))


def replay_pushes(buf, g, lo, hi):
    for i in range(lo, hi):
        buf.push(g["push_s"][i], g["push_a"][i], g["push_r"][i], g["push_ns"][i], g["push_d"][i])


@pytest.mark.parametrize("algo,case", PER_CASES)
def test_per_buffer_oracle_matches_reference_fixture(algo, case):
    g = load(f"per_{algo}_{case}")
    m = per_meta(g)
    buf = OP.PERBufferOracle(m["max_len"], m["alpha"])
    for si in range(len(g["steps"])):
        replay_pushes(buf, g, *per_pushes(g, si))
        before = g[f"s{si}_prio_before"]
        assert len(buf) == before.shape[0]
        # free-running priorities stay within the powf ulp of the reference's; then teacher-force them
        assert ulp_diff_f32(np.array(buf.priorities, np.float32), before).max() <= 2
        buf.priorities = type(buf.priorities)(before.tolist(), maxlen=m["max_len"])
        u = g[f"s{si}_u"]
        buf._uniforms = lambda b, _u=u: _u
        s, a, r, ns, d, w, idx = buf.sample(m["B"], float(g[f"s{si}_beta"]))
        assert np.array_equal(idx, g[f"s{si}_idx"])
        assert np.array_equal(bits(OP.normalised_priorities(before)), bits(g[f"s{si}_P"]))
        for got, key in zip((s, a, r, ns, d), ("s", "a", "r", "ns", "d")):
            assert np.array_equal(bits(got), bits(g[f"s{si}_batch_{key}"])), key
        assert ulp_diff_f32(w, g[f"s{si}_w"]).max() <= 2
        buf.update_priorities(idx, g[f"s{si}_td"])
        assert ulp_diff_f32(np.array(buf.priorities, np.float32), g[f"s{si}_prio_after"]).max() <= 2


@pytest.mark.parametrize("algo,case", PER_CASES)
def test_weighted_update_oracles_match_reference_fixture(algo, case):
    g = load(f"per_{algo}_{case}")
    m = per_meta(g)
    orc, step_fn = per_oracle(algo, g)
    steps = [int(x) for x in g["steps"]]
    for si, step in enumerate(steps):
        b = [g[f"s{si}_batch_{k}"] for k in ("s", "a", "r", "ns", "d")]
        normals = [g[f"s{si}_normal{j}"] for j in range(int(g[f"s{si}_n_normal"]))]
        info = step_fn(step, b, g[f"s{si}_w"], normals)
        ref = g[f"s{si}_info"]
        assert len(info) == len(ref)
        tdp = per_td_position(algo, len(info))
        td = np.asarray(info[tdp], np.float32).reshape(-1, 1)
        np.testing.assert_allclose(td, g[f"s{si}_td"], rtol=2e-5, atol=2e-6)
        flat = [float(np.mean(x)) if i == tdp else float(x) for i, x in enumerate(info)]
        np.testing.assert_allclose(np.array(flat), ref, rtol=2e-5, atol=1e-6)
    si, n, lr = len(steps) - 1, len(steps), m["lr"]
    if f"s{si}_actor.base_net.0.weight" not in g.files:
        return
    if algo in ("ddpg", "td3"):
        tags = ["actor", "target_actor"] + (["critic", "target_critic"] if algo == "ddpg" else
                                            ["critic_1", "critic_2", "target_critic_1", "target_critic_2"])
        for tag in tags:
            for (w, b_), (rw, rb) in zip(getattr(orc, tag), ddpg_params_from_golden(g, si, tag)):
                assert weights_close(w, rw, lr, n) and weights_close(b_, rb, lr, n), tag
    else:
        ref = sac_params_from_golden(g, si, "actor")
        assert_sac_actor_close(orc.actor, ref["params"], lr, n)
        for i in (0, orc.n - 1):
            tag = f"critic_{i + 1}" if algo == "sac" else f"critic_{i}"
            for (w, b_), (rw, rb) in zip(orc.critics[i], sac_params_from_golden(g, si, tag)):
                assert weights_close(w, rw, lr, n) and weights_close(b_, rb, lr, n), tag


def test_restatements_hold_on_random_shapes():
    """Property check (hypothesis): any length, any float32 priorities -- the restated pairwise sum equals
    ndarray.sum() bit for bit and the restated draw equals RandomState.choice."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    @settings(max_examples=60, deadline=None)
    @given(n=st.integers(1, 700), seed=st.integers(0, 2 ** 31 - 1), spread=st.integers(0, 12))
    def check(n, seed, spread):
        rng = np.random.default_rng(seed)
        prio = (10.0 ** rng.uniform(-spread, 1, n)).astype(np.float32)
        assert bits(OP.pairwise_sum_f32(prio)) == bits(prio.sum())
        P = OP.normalised_priorities(prio)
        ref = prio.copy()
        ref /= ref.sum()
        assert np.array_equal(bits(P), bits(ref))
        B = int(rng.integers(1, 33))
        np.random.seed(seed % (2 ** 32))
        state = np.random.get_state()
        want = np.random.choice(n, B, p=ref)
        np.random.set_state(state)
        assert np.array_equal(OP.choice_indices(P, np.random.random_sample(B)), want)

    check()
