import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")
    # The stated tolerances are those of the default operand type (fp16).  Tests that cover bf16 select it themselves
    # (monkeypatch / explicit dtype arguments), so an inherited DUCOSY_PRECISION must not silently change what is tested.
    os.environ["DUCOSY_PRECISION"] = "fp16"


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
