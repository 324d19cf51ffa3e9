import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return np.load(GOLDEN / f"{name}.npz", allow_pickle=False)


@pytest.fixture(scope="session")
def golden():
    return load_golden


def group_keys(npz, prefix_sep="/"):
    """{'case': {'field': array}} from flat 'case/field' keys."""
    out = {}
    for k in npz.files:
        case, _, field = k.partition(prefix_sep)
        out.setdefault(case, {})[field] = npz[k]
    return out
