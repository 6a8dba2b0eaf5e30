"""CPU oracle for the learner hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product package never does.
See ``oracle/oracle.py`` for the per-function reference citations.
"""
