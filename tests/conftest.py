import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


@pytest.fixture
def golden():
    return load_golden


def rel_l2(a, b):
    """||a-b||_2 / ||b||_2 over the whole tensor (the tolerance metric DESIGN.md states)."""
    import torch
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def record_parity(name, **values):
    """Append measured parity numbers (rel-L2, SNR, bit-equality) of a GPU test to a JSON file that travels back from the GPU
    box (gpurun_out/parity_measured.json); profiles/r2_parity.json is a committed copy of one such run."""
    import json
    path = os.environ.get("ADB_PARITY_OUT", os.path.join(ROOT, "gpurun_out", "parity_measured.json"))
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        data = {}
        if os.path.exists(path):
            with open(path) as f:
                data = json.load(f)
        data[name] = {k: (float(v) if isinstance(v, (int, float)) and not isinstance(v, bool) else v) for k, v in values.items()}
        with open(path, "w") as f:
            json.dump(data, f, indent=1, sort_keys=True)
    except OSError:
        pass
