"""CPU oracle for the FRUITS ISS + sieve hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``fruits_b200/`` may import this
package.  Allowed importers: ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- and there only
as the checker or as the CPU baseline that is reported beside the GPU number.

The oracle restates the reference algorithm (irkri/fruits 1.0.0, pure Python
+ numba) in plain C (``fruits_oracle.c``, numeric kernels) and numpy
(``pipeline.py``, the fit/transform orchestration, word enumeration, cache
plan, RNG consumption order).  Each function cites the reference file:line it
follows.

Parity is PINNED: ``oracle/gen_golden.py`` imports the real reference in the
build container (``/root/reference`` with the ``np.NINF`` shim), checks this
oracle against it on the reference's own known-answer tests and on seeded
random inputs, and freezes the reference outputs under ``tests/golden/``.
``tests/test_oracle_golden.py`` re-checks the oracle against those frozen
vectors without needing the reference.
"""
from .build import build_oracle, load_oracle  # noqa: F401
