"""GPU parity: running normaliser (csrc/normalizer.cu) through the C ABI versus the oracle
and the fixtures dumped from the reference's RunningNormalizer (src/utils.py:68-117).

Tolerance: the state is float64 on both sides; the device reduces a batch in a different
(fixed) order than numpy's pairwise mean/var, so statistics agree to rel 1e-12 -- far inside
the north-star's fp32 rel 1e-5."""
import numpy as np
import pytest

from oracle import her as OH
from tests.helpers import load

pytestmark = pytest.mark.gpu

RTOL = 1e-12


@pytest.mark.parametrize("tag,dim", [("obs", 19), ("dg", 3)])
def test_update_and_normalize_match_reference_fixture(tag, dim):
    from gcrl_b200 import RunningNormalizer
    g = load("normalizer")
    nz = RunningNormalizer(size=dim)
    assert nz.count == 1e-8 and np.all(nz.mean == 0) and np.all(nz.var == 1)
    for i in range(6):
        nz.update(g[f"{tag}_x{i}"])
        np.testing.assert_allclose(nz.mean, g[f"{tag}_mean{i}"], rtol=RTOL, atol=1e-14)
        np.testing.assert_allclose(nz.var, g[f"{tag}_var{i}"], rtol=RTOL, atol=1e-14)
        assert nz.count == pytest.approx(float(g[f"{tag}_count{i}"]), rel=1e-15)
    out = nz.normalize(g[f"{tag}_q"])
    assert out.dtype == np.float64 and out.shape == g[f"{tag}_qn"].shape
    np.testing.assert_allclose(out, g[f"{tag}_qn"], rtol=1e-11, atol=1e-13)
    assert np.abs(out).max() <= 5.0 and (np.abs(out) == 5.0).any()     # clip is exercised


@pytest.mark.parametrize("n,dim,dtype", [(1, 3, np.float32), (2, 7, np.float64), (257, 19, np.float64),
                                         (100_000, 20, np.float32), (1_000_003, 3, np.float32)])
def test_large_and_ragged_batches_against_oracle(n, dim, dtype):
    from gcrl_b200 import RunningNormalizer
    rng = np.random.default_rng(n)
    nz, orc = RunningNormalizer(size=dim), OH.RunningNormalizerOracle(dim)
    for _ in range(3):
        x = (rng.standard_normal((n, dim)) * rng.uniform(0.1, 3, dim) + rng.uniform(-2, 2, dim)).astype(dtype)
        nz.update(x)
        orc.update(x)
    np.testing.assert_allclose(nz.mean, orc.mean, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(nz.var, orc.var, rtol=1e-10, atol=1e-12)
    assert nz.count == pytest.approx(orc.count, rel=1e-15)
    q = (rng.standard_normal((min(n, 4096), dim)) * 6).astype(dtype)
    np.testing.assert_allclose(nz.normalize(q), orc.normalize(q), rtol=1e-10, atol=1e-12)


def test_yaml_round_trip_and_float32_narrowing(tmp_path):
    """save/load keep the reference's YAML keys; load narrows to float32 (src/utils.py:114-115)."""
    import yaml
    from gcrl_b200 import RunningNormalizer
    rng = np.random.default_rng(5)
    nz = RunningNormalizer(size=4)
    nz.update(rng.standard_normal((50, 4)))
    path = str(tmp_path / "norm" / "obs.yaml")
    nz.save(path)
    data = yaml.safe_load(open(path))
    assert sorted(data) == ["clip_range", "count", "mean", "var"]
    nz2 = RunningNormalizer(size=4)
    nz2.load(path)
    np.testing.assert_array_equal(nz2.mean, np.array(data["mean"], np.float32).astype(np.float64))
    np.testing.assert_array_equal(nz2.var, np.array(data["var"], np.float32).astype(np.float64))
    assert nz2.count == data["count"]


def test_agent_glue_concat_matches_oracle():
    """normalize_state_batch = concat([norm(obs), norm(dg)]) (src/agent.py:1434-1445)."""
    from gcrl_b200 import RunningNormalizer
    rng = np.random.default_rng(9)
    obs, dg = rng.standard_normal((8, 19)), rng.standard_normal((8, 3)).astype(np.float32)
    no, ng = RunningNormalizer(19), RunningNormalizer(3)
    oo, og = OH.RunningNormalizerOracle(19), OH.RunningNormalizerOracle(3)
    for a, b in ((no, oo), (ng, og)):
        x = rng.standard_normal((16, a.size))
        a.update(x)
        b.update(x)
    got = np.concatenate([no.normalize(obs), ng.normalize(dg)], -1)
    want = np.concatenate([oo.normalize(obs), og.normalize(dg)], -1)
    np.testing.assert_allclose(got, want, rtol=1e-11, atol=1e-13)
