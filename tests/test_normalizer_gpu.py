"""GPU parity: running normaliser (csrc/normalizer.cu) through the C ABI versus the oracle
and the fixtures dumped from the reference's RunningNormalizer (src/utils.py:68-117).

Bar: for reference-scale batches (<= 8192 rows, dim >= 2 -- src/env.py:165-172 feeds 2-4 rows per
env per step) the device walks the rows in numpy's order, in the input dtype, with unfused IEEE
operations, so the float64 running statistics and normalised outputs are BIT-IDENTICAL to the
reference's.  Larger batches use a parallel float64 Welford/Chan reduction: it agrees with a
float64 evaluation to rel 1e-10 and with the reference's float32-accumulated moments of float32
inputs to the reference's own rounding error (~sqrt(n) * 6e-8, stated per test)."""
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
        np.testing.assert_array_equal(nz.mean, g[f"{tag}_mean{i}"])          # bit-exact float64
        np.testing.assert_array_equal(nz.var, g[f"{tag}_var{i}"])
        assert nz.count == float(g[f"{tag}_count{i}"])
    out = nz.normalize(g[f"{tag}_q"])
    assert out.dtype == np.float64 and out.shape == g[f"{tag}_qn"].shape
    np.testing.assert_array_equal(out, g[f"{tag}_qn"])
    assert np.abs(out).max() <= 5.0 and (np.abs(out) == 5.0).any()     # clip is exercised


@pytest.mark.parametrize("n,dim,dtype", [(1, 3, np.float32), (2, 7, np.float64), (257, 19, np.float64),
                                         (8192, 20, np.float32), (4000, 3, np.float32)])
def test_reference_scale_batches_bit_exact(n, dim, dtype):
    from gcrl_b200 import RunningNormalizer
    rng = np.random.default_rng(n)
    nz, orc = RunningNormalizer(size=dim), OH.RunningNormalizerOracle(dim)
    for _ in range(3):
        x = (rng.standard_normal((n, dim)) * rng.uniform(0.1, 3, dim) + rng.uniform(-2, 2, dim)).astype(dtype)
        nz.update(x)
        orc.update(x)
        np.testing.assert_array_equal(nz.mean, orc.mean)
        np.testing.assert_array_equal(nz.var, orc.var)
        assert nz.count == orc.count
    q = (rng.standard_normal((min(n, 4096), dim)) * 6).astype(dtype)
    np.testing.assert_array_equal(nz.normalize(q), orc.normalize(q))


@pytest.mark.parametrize("n,dim,dtype", [(100_000, 20, np.float32), (1_000_003, 3, np.float32),
                                         (50_000, 19, np.float64), (9000, 1, np.float32)])
def test_large_batches_parallel_reduction(n, dim, dtype):
    from gcrl_b200 import RunningNormalizer
    rng = np.random.default_rng(n)
    nz = RunningNormalizer(size=dim)
    exact, ref = OH.RunningNormalizerOracle(dim), OH.RunningNormalizerOracle(dim)
    for _ in range(3):
        x = (rng.standard_normal((n, dim)) * rng.uniform(0.1, 3, dim) + rng.uniform(-2, 2, dim)).astype(dtype)
        nz.update(x)
        exact.update(x.astype(np.float64))     # the same moments evaluated in float64
        ref.update(x)                          # numpy accumulates float32 inputs in float32
    np.testing.assert_allclose(nz.mean, exact.mean, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(nz.var, exact.var, rtol=1e-10, atol=1e-12)
    assert nz.count == pytest.approx(exact.count, rel=1e-15)
    # against the reference's own float32 accumulation: its rounding error, ~sqrt(n) * 6e-8 * |x|
    tol = 8 * np.sqrt(n) * 6e-8 if dtype == np.float32 else 1e-10
    np.testing.assert_allclose(nz.mean, ref.mean, rtol=tol, atol=tol * 3)
    np.testing.assert_allclose(nz.var, ref.var, rtol=tol * 4, atol=tol * 10)
    q = (rng.standard_normal((4096, dim)) * 6).astype(dtype)
    np.testing.assert_allclose(nz.normalize(q), exact.normalize(q), rtol=1e-9, atol=1e-10)


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
    np.testing.assert_array_equal(got, want)


@pytest.mark.parametrize("world,dtype,tol", [(2, np.float64, 1e-12), (4, np.float64, 1e-12), (8, np.float32, 2e-6)])
def test_data_parallel_update_equals_single_process_on_the_concatenated_batch(world, dtype, tol):
    """every rank contributes the moments of its own envs' observations; all ranks fold all moments
    in rank order.  N ranks on per-rank batches == one process on the concatenated batch (the oracle, i.e. the
    reference's update) to float64 rounding -- float32 batches to float32 rounding, because NumPy forms
    batch_var * n in the batch dtype (src/utils.py:88) -- and the replicas are bit-identical."""
    from gcrl_b200 import RunningNormalizer
    rng = np.random.default_rng(world)
    dim = 19
    ranks = [RunningNormalizer(dim) for _ in range(world)]
    single = OH.RunningNormalizerOracle(dim)
    for step in range(5):
        parts = [(rng.standard_normal((3 + r, dim)) * (1 + step) + 0.3 * r).astype(dtype) for r in range(world)]
        gathered = np.stack([nz.batch_moments(x) for nz, x in zip(ranks, parts)])      # the all-gather
        for nz, x in zip(ranks, parts):
            nz.enable_data_parallel(gather=lambda m, _g=gathered: _g)
            nz.update(x)
        single.update(np.concatenate(parts, 0))
        np.testing.assert_allclose(ranks[0].mean, single.mean, rtol=tol, atol=tol)
        np.testing.assert_allclose(ranks[0].var, single.var, rtol=tol, atol=tol)
        assert ranks[0].count == pytest.approx(single.count, rel=1e-15)
        for nz in ranks[1:]:
            assert np.array_equal(nz.mean, ranks[0].mean) and np.array_equal(nz.var, ranks[0].var)
    q = rng.standard_normal((7, dim))
    np.testing.assert_allclose(ranks[-1].normalize(q), single.normalize(q), rtol=100 * tol, atol=100 * tol)
