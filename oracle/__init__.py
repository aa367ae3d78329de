"""CPU oracle for the HybridFusion hot path and the ECE binning.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker.  The product
path (the package next to this directory) never imports from here and fails
loudly when its CUDA library is missing.

Parity pin: the restatements in this package are checked against golden
vectors produced by importing the *unmodified* reference modules from
``/root/reference/src`` (``oracle/make_golden.py`` -> ``tests/golden/*.npz``),
see ``tests/test_oracle_golden.py``.
"""
