"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the hilbert-quantization hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import it, and only as the checker.  The product package
(``hilbert_quantization_b200``) never imports this package and fails loudly
when its CUDA library is missing.

Parity status: PINNED.  ``oracle/pin_against_reference.py`` (run in the
authoring container, where ``/root/reference`` exists) checks every function
of ``oracle.hilbert_oracle`` against the reference's own classes, and
``tests/golden/make_golden.py`` freezes reference outputs as ``.npz``
fixtures that ``tests/test_oracle_golden.py`` re-checks on every run.
"""
