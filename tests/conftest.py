import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built_artifacts():
    """Everything the suite loads is built in-tree once per session (all are cheap, CPU-only
    builds except libsqoa_b200.so, which nvcc cross-compiles in ~20 s when it is missing)."""
    import oracle

    oracle.build(with_reference=True)
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")], check=True)
    lib = os.path.join(ROOT, "seqoia_b200", "libsqoa_b200.so")
    synth = os.path.join(ROOT, "seqoia_b200", "libsqoa_synth.so")
    if not (os.path.exists(lib) and os.path.exists(synth)):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "seqoia_b200", "csrc")], check=True)
    yield
