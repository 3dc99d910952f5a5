"""CPU oracle for the SHMGAN hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  Nothing under ``shmgan_b200/``
imports it; the product path has no CPU fallback.

PARITY UNPINNED: TensorFlow / Keras / tensorflow-addons are not installable in this
image and the reference ships no golden vectors, so the restatement is pinned only by
the known-answer tests derived from ``*_summary.txt`` (param counts, layer shapes) and
hand-checkable padding impulses (see ``tests/test_oracle_kat.py``).
"""
from .shmgan_oracle import *  # noqa: F401,F403
