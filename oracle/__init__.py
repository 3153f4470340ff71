"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the IISAN(Cached) training hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and only as the checker
(or as the timed CPU arm), never as the thing that is shipped.  The product path
(``iisan_b200``) must never import this package.

Parity status: PINNED.  ``oracle/make_golden.py`` runs the reference's own ``model`` package
(``/root/reference/Code_Cached/model`` and ``Code_Cached_Asym/model``) on seeded synthetic inputs in
the build container and freezes its outputs under ``tests/golden/``; ``tests/test_oracle_golden.py``
checks this restatement against those fixtures on every run.  (The reference ships no golden
vectors or tests of its own -- SURVEY.md section 4.)
"""
