import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _build_oracle_port():
    """The plain-C oracle is test infrastructure: (re)build it when stale."""
    so = ROOT / "oracle" / "liboracle.so"
    src = ROOT / "oracle" / "oracle.c"
    if not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        env = dict(os.environ)
        env.pop("CC", None)
        subprocess.run(["make", "-C", str(ROOT / "oracle"), "port"], check=True, env=env,
                       stdout=subprocess.DEVNULL)
    yield


@pytest.fixture(autouse=True)
def _restore_library_options():
    yield
    import util
    util.reset_options()
