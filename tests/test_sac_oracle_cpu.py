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
