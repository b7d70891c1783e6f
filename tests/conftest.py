import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
TESTS = os.path.dirname(os.path.abspath(__file__))
if TESTS not in sys.path:
    sys.path.insert(0, TESTS)
GOLDEN = os.path.join(TESTS, "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def record_parity(name, **vals):
    """Append the OBSERVED error of a parity comparison to gpurun_out/parity_observed.jsonl (when that scratch directory
    exists): the tolerances asserted in the tests are ~10x these and are re-derived from this log when kernels change."""
    import json
    d = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "parity_observed.jsonl"), "a") as f:
            f.write(json.dumps({"test": name, **{k: float(v) for k, v in vals.items()}}) + "\n")
