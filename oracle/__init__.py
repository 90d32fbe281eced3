"""TEST INFRASTRUCTURE ONLY — CPU oracle for the VAESNe hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it, and only as the checker (or as the timed CPU
baseline).  The product path (``vaesne-dev_b200/``) never imports this package
and raises if its CUDA library is missing.

Parity status: PINNED.  ``oracle/make_golden.py`` imports the real reference
(``/root/reference/package/VAESNe``) in the build container, runs it on seeded
synthetic inputs with dropout 0 and recorded reparameterisation noise, and
writes ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks this
restatement against those vectors on every CPU test run.
"""
