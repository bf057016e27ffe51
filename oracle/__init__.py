"""CPU oracle for the mliis inner-loop adaptation hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``mliis_b200/`` may import this package;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it, and only as the checker / the timed CPU baseline.

PARITY UNPINNED: the reference (ml4ai/mliis) is TensorFlow-1.15 graph code that
cannot be imported in this environment and ships no tests, golden vectors or
fixtures for this path.  The oracle is therefore a restatement
of the reference graph from its source (file:line cited per function) plus the
documented TF-1.15 op semantics, self-checked by finite differences and by
cross-checks against independent torch primitives (see tests/test_oracle.py).
"""
