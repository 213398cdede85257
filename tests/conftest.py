import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_sessionstart(session):
    """The tests load ``libfruits_b200.so``; build it if the tree is fresh
    (``__graft_entry__.build()`` does the same; nvcc needs no GPU)."""
    from fruits_b200 import _backend, build
    if not os.path.exists(_backend.LIB_PATH) and os.path.exists(build.NVCC):
        build.build()


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def pytest_sessionfinish(session, exitstatus):
    """Parity reports of the weighted configurations (rtol violations, count
    flips) next to the other run artefacts."""
    import json
    import helpers
    out = os.path.join(ROOT, "gpurun_out")
    if helpers.REPORTS and os.path.isdir(out):
        with open(os.path.join(out, "parity_report.json"), "w") as f:
            json.dump(helpers.REPORTS, f, indent=1, sort_keys=True)
