"""ORACLE — test infrastructure only.

CPU restatements (torch / numpy) of the reference algorithms on the hot path, each function citing the
reference file:line it follows, pinned to the third-party classes the reference calls and to golden outputs
of the unmodified reference code (tests/golden/).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s CPU-baseline / ``--impl reference`` legs may import this package — as the checker or the
timed CPU reference, never as part of the product: ``mertools_b200/`` does not import it
(tests/test_host_logic.py enforces that) and has no CPU fallback.

Pinning status: every restatement is checked against outputs of the unmodified reference code or of the
third-party class the reference calls (tests/test_oracle.py, tests/test_golden.py).  One pin is weaker than the
others: ``encoders.vggish_embeddings`` -- TensorFlow / tf_slim are not installed, so the unmodified reference graph
definition and extractor (vggish_slim.py, extract_vggish_embedding.py) were run over a torch-backed stand-in for the
TF calls they make (tests/golden/tf_slim_shim.py -> vggish_golden.npz): structure, variable names and extractor logic
are the reference's, TensorFlow's float arithmetic is not exercised.
"""
