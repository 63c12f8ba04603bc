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


def pytest_sessionstart(session):
    """The library is a build artefact (git-ignored): compile it when the tree has none yet and nvcc is here, so that a fresh
    checkout can run the suite without a separate build step.  Nothing is built when the .so exists (the GPU box gets it
    with the snapshot) and nothing is faked when it cannot be built -- the ABI tests then fail loudly."""
    import shutil
    lib = os.path.join(ROOT, "ducosy_gan_b200", "lib", "libducosy_sm100.so")
    if not os.path.exists(lib) and (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")):
        from ducosy_gan_b200 import build as _build
        _build.build()


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
