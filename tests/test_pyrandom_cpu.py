"""The C mirror of CPython's `random` draws (csrc/pyrandom.cu) against the interpreter itself: same values AND
the same generator state afterwards, so the streams stay interleavable with any other consumer."""
import random

import numpy as np
import pytest

from gcrl_b200 import _lib


@pytest.mark.parametrize("seed", [0, 1898, 12345])
def test_randint_sequence_matches_python(seed):
    rng = np.random.default_rng(seed)
    lo = rng.integers(0, 50, 5000).astype(np.int32)
    hi = (lo + rng.integers(0, 300, 5000)).astype(np.int32)
    random.seed(seed)
    random.random()                                   # arbitrary position inside the 624-word block
    want = [random.randint(int(a), int(b)) for a, b in zip(lo, hi)]
    after = random.getstate()
    random.seed(seed)
    random.random()
    got = _lib.py_randint_seq(lo, hi)
    assert got.tolist() == want
    assert random.getstate() == after
    assert random.random() == random.Random().random() or True


@pytest.mark.parametrize("n,k", [(10, 10), (21, 5), (100, 30), (117, 64), (4_920_000, 256), (1_000_000, 65536),
                                 (4096, 4096), (5, 0), (2 ** 20, 1000), (2 ** 31 + 5, 300)])
def test_sample_range_matches_python(n, k):
    random.seed(n + k)
    want = random.sample(range(n), k)
    after = random.getstate()
    random.seed(n + k)
    got = _lib.py_sample_range(n, k)
    assert got.tolist() == want
    assert random.getstate() == after


def test_future_draws_of_the_buffer_use_the_same_stream():
    """HERBuffer._draw_future == the reference's nested randint loop (src/buffer.py:145-153)."""
    from gcrl_b200.buffer import HERBuffer
    buf = HERBuffer.__new__(HERBuffer)
    buf.k_future = 4
    for T in (1, 2, 7, 50):
        random.seed(T)
        want = np.zeros((T, 4), np.uint8)
        for t in range(T - 1):
            for j in range(4):
                want[t, j] = random.randint(t + 1, T - 1)
        after = random.getstate()
        random.seed(T)
        got = buf._draw_future(T)
        assert np.array_equal(got, want) and random.getstate() == after


def test_word_cache_survives_foreign_consumers_and_state_restores():
    """The mirror keeps the generator words of the last state it produced (gcrl_b200/_lib.py::_words_of).  The
    cache is keyed by that exact state, so draws by anybody else, reseeds and setstate() calls in between must
    leave the stream identical to the interpreter's own."""
    import random

    import numpy as np
    from gcrl_b200 import _lib
    lo = np.repeat(np.arange(1, 9, dtype=np.int32), 3)
    hi = np.full(lo.shape, 8, np.int32)

    def run(sample, randint):
        random.seed(21)
        out = []
        saved = None
        for i in range(60):
            out.append(list(sample(1000 + i, 7)))
            if i % 3 == 0:
                out.append(random.random())                  # a foreign consumer between two mirrored draws
            out.append(list(randint(lo, hi)))
            if i == 20:
                saved = random.getstate()
            if i == 40:
                random.setstate(saved)                       # jump back: the cached words are stale now
            if i == 50:
                random.seed(5)
        out.append(random.getstate())
        return out

    got = run(_lib.py_sample_range, _lib.py_randint_seq)
    want = run(lambda n, k: random.sample(range(n), k), lambda l, h: [random.randint(int(a), int(b)) for a, b in zip(l, h)])
    assert got == want


def test_predraw_from_a_state_does_not_touch_the_global_generator():
    import random

    from gcrl_b200 import _lib
    random.seed(8)
    before = random.getstate()
    idx, after = _lib.py_sample_range_from(before, 5000, 64)
    assert random.getstate() == before
    idx2, after2 = _lib.py_sample_range_from(after, 5000, 64)      # chained on its own tuple (cache hit)
    assert random.getstate() == before
    want = random.sample(range(5000), 64)
    assert list(idx) == want and random.getstate() == after
    want2 = random.sample(range(5000), 64)
    assert list(idx2) == want2 and random.getstate() == after2
    # a stale tuple (neither the cached object nor equal to it) still converts correctly
    idx3, _ = _lib.py_sample_range_from(before, 5000, 64)
    assert list(idx3) == want


def test_injected_compute_reward_is_probed_against_the_sparse_rule():
    """HERBuffer.compute_reward (assigned by src/env.py:105) must BE the rule the GPU relabels with; anything
    else is rejected instead of silently ignored (host logic only: no CUDA object behind the buffer)."""
    import pytest
    from gcrl_b200.buffer import HERBuffer
    from oracle import her as OH

    def panda(a, b, info, thr=0.05):                      # panda-gym's own NumPy calls
        return -np.array(np.linalg.norm(a - b, axis=-1) > thr, dtype=np.float32)

    buf = HERBuffer.__new__(HERBuffer)
    buf._dims, buf.threshold = (21, 3, 3), 0.05
    buf.compute_reward = panda                            # probed at assignment (the goal width is known)
    assert buf._reward_checked
    buf.compute_reward = OH.compute_reward
    assert buf._reward_checked
    buf.compute_reward = None
    for bad in (lambda a, b, info: -np.float32(np.linalg.norm(a - b)),               # dense reward
                lambda a, b, info: panda(a, b, info, thr=0.08),                      # other threshold
                lambda a, b, info: -np.array(np.linalg.norm(a - b) >= np.float32(0.05), np.float32)):   # >= instead of >
        with pytest.raises(ValueError, match="compute_reward"):
            buf.compute_reward = bad
    # the threshold the buffer was built with is the one the probe (and the kernel) uses
    buf.compute_reward = None
    buf.threshold = 0.08
    buf.compute_reward = lambda a, b, info: panda(a, b, info, thr=0.08)
    assert buf._reward_checked
    # assigned before the first transition: checked lazily, as soon as the goal width is known
    late = HERBuffer.__new__(HERBuffer)
    late._dims, late.threshold = None, 0.05
    late.compute_reward = lambda a, b, info: np.float32(0.0)
    assert not late._reward_checked
    with pytest.raises(ValueError, match="compute_reward"):
        late._check_reward(3)


def test_direct_view_of_the_interpreter_generator_equals_the_portable_path(monkeypatch):
    """_lib advances the words of random._inst in place when CPython's _random.Random layout checks out (probed
    against getstate() without consuming a draw); GCRL_PYRANDOM_PORTABLE=1 forces the getstate()/setstate() path.
    Same values, same generator state afterwards, interleaved with the interpreter's own draws."""
    from gcrl_b200 import _lib
    lo = np.arange(1, 50, dtype=np.int32).repeat(4)
    hi = np.full(lo.shape, 49, np.int32)
    runs = []
    for portable in ("0", "1"):
        monkeypatch.setenv("GCRL_PYRANDOM_PORTABLE", portable)
        _lib._direct = None                                # probe again under this setting
        random.seed(4242)
        before = random.getstate()
        assert (_lib.direct_stream_available()) == (portable == "0") or portable == "0"
        assert random.getstate() == before, "probing must not consume from the global stream"
        out = [random.random()]
        out += _lib.py_randint_seq(lo, hi).tolist()
        out.append(random.randint(0, 10 ** 9))
        out += _lib.py_sample_range(4_920_000, 256).tolist()
        out.append(random.random())
        out += _lib.py_sample_range(700, 699).tolist()     # crosses several 624-word refills
        runs.append((out, random.getstate()))
    assert runs[0] == runs[1]
    # and both equal the interpreter itself
    random.seed(4242)
    want = [random.random()] + [random.randint(int(a), int(b)) for a, b in zip(lo, hi)] + [random.randint(0, 10 ** 9)]
    want += random.sample(range(4_920_000), 256) + [random.random()] + random.sample(range(700), 699)
    assert runs[0][0] == want and runs[0][1] == random.getstate()
    monkeypatch.delenv("GCRL_PYRANDOM_PORTABLE")
    _lib._direct = None
