"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the FEABAS xcorr hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` leg may import it, and only as the checker (or as the
timed CPU baseline), never as the thing shipped.  ``feabas_b200`` never
imports this package.

Parity status: the reference ships no tests or golden vectors,
so the oracle is pinned against outputs of the *unmodified* reference functions
executed in the build container (``oracle/ref_loader.py`` +
``oracle/make_golden.py`` -> ``tests/golden/*.npz``).
"""
