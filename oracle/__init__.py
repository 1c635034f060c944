"""ORACLE — test infrastructure only.

CPU restatements (torch / numpy) of the reference algorithms on the hot path, each function citing the
reference file:line it follows, pinned to the third-party classes the reference calls and to golden outputs
of the unmodified reference code (tests/golden/).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs may import this package — as the checker or the
timed CPU reference, never as part of the product: ``mertools_b200/`` does not import it
(tests/test_host_logic.py enforces that) and has no CPU fallback.

Pinning status: every restatement is checked against outputs of the unmodified reference code or of the
third-party class the reference calls (tests/test_oracle.py, tests/test_golden.py), with one exception —
``encoders.vggish_embeddings`` is PARITY UNPINNED: the reference graph needs TensorFlow / tf_slim, which are not
installed, so it follows the definition file only (cross-checked against the torchvggish form of the network).
"""
