"""CPU oracle for the HER-sample + off-policy-update hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported, linked or
executed by the product (``goal-conditioned-rl-framework_b200/``).  The only
callers are ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` -- and there only as the checker or
as the timed CPU baseline, never as the thing shipped.

Every function restates, in plain NumPy / pure Python, the algorithm of the
reference file:line it cites (paths are relative to the upstream repository
CodeKnight314/Goal-Conditioned-RL-Framework).

Parity pinning (see ``tests/golden/make_golden.py``):
  * HER store / relabel / sample, RunningNormalizer and the DDPG / TD3 update
    are pinned against the *unmodified reference classes* executed in the build
    container; the resulting inputs / RNG draws / outputs are committed under
    ``tests/golden/*.npz`` and checked by ``tests/test_oracle_golden.py``.
  * ``compute_reward`` lives in the un-vendored, un-pinned third-party package
    ``panda-gym`` (requirements.txt:10).  It is restated from the published
    sparse rule ``-(||achieved - desired||_2 > 0.05)`` (float32): the reference
    has no test or vector that pins it, so for this one function parity is
    **unpinned** beyond the restated rule and its known-answer edge cases.
"""
