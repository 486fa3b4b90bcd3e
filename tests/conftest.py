import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        z = np.load(GOLDEN / f"{name}.npz")
        return {k: torch.from_numpy(z[k]) if z[k].ndim else z[k].item() for k in z.files}
    return load


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the checker (C oracle) once; the product library is built by __graft_entry__.build()."""
    import oracle
    oracle.build()
