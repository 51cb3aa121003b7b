"""CPU oracle for the veomni_b200 hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``veomni_b200/`` may import this package: the only
legitimate users are ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` (as the checker / the timed CPU baseline, never as the
product path).

Each function restates, in plain torch-on-CPU / numpy, the algorithm of one reference function on
the path and cites the reference ``file:line`` it follows (paths relative to the VeOmni tree).
The restatement is pinned against the reference itself: ``tests/golden/make_golden.py`` imports
the reference from ``/root/reference`` (available only in the authoring container), runs the
reference function and this oracle on the same seeded inputs, asserts agreement, and stores the
inputs/outputs as fixtures under ``tests/golden/`` — ``tests/test_oracle_golden.py`` re-checks
the oracle against those fixtures wherever the tests run (the GPU box has no reference tree).

Pieces whose arithmetic lives in third-party CUDA/Triton code that cannot run here are restated
from their published algorithm and pinned on the reference's own call sites instead; each such
function says "parity unpinned" or names the reference test that pins it.
"""
