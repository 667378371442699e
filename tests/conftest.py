import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def _cuda_devices():
    """Number of usable CUDA devices as the product library sees them (0 if it cannot even be built/loaded)."""
    try:
        from pyaudiorestoration_b200 import _lib
        return int(_lib.lib().par_device_count())
    except Exception:                                            # noqa: BLE001
        return 0


def pytest_collection_modifyitems(config, items):
    """Without a GPU the `gpu` tests are SKIPPED (not errors): a plain `pytest tests` is green on a CPU box and runs
    everything on the B200 box.  The product itself never falls back -- it raises."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if gpu_items and _cuda_devices() <= 0:
        skip = pytest.mark.skip(reason="no CUDA device: GPU parity tests run on the B200 box (gpurun)")
        for it in gpu_items:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
