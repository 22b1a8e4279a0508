import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_scene():
    return dict(np.load(GOLDEN / "golden_scene.npz"))


@pytest.fixture(scope="session")
def golden_aggregate():
    return dict(np.load(GOLDEN / "golden_aggregate.npz"))


@pytest.fixture(scope="session")
def golden_render():
    return dict(np.load(GOLDEN / "golden_render.npz"))
