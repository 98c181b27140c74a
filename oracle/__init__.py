"""ORACLE — test infrastructure only.

CPU restatement of the reference's aggregation-AMG hot path (nicknytko/ml-amg).
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this package; the product path
(`ml-amg_b200/`) never does.

Parity status
-------------
* reference-owned scipy expressions (`oracle.reference_path`): pinned against the
  UNMODIFIED reference modules imported in the build container
  (`tests/golden/make_golden.py` -> `tests/golden/*.npz`).
* pyamg routines (`oracle.pyamg_restated`, `oracle/amg_core_restated.c`):
  **PARITY UNPINNED** — pyamg is an un-vendored, un-pinned, un-installable
  third-party dependency; its 4.x algorithm is restated from its published
  source and cross-checked only against an independent pure-Python restatement.
* multilevel cycle / PCG / L1-Jacobi (`oracle.multilevel`): **PARITY UNPINNED** —
  no implementation exists in the reference.
"""
from . import pyamg_restated, reference_path, multilevel  # noqa: F401
