import json
import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def load_golden(name: str):
    z = np.load(GOLDEN / f"{name}.npz", allow_pickle=False)
    cfg = json.loads(str(z["cfg_json"]))
    return z, cfg


@pytest.fixture(scope="session")
def golden():
    return load_golden
