import os
import sys
import warnings

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore", category=FutureWarning)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def load_golden(name):
    import torch
    return torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)


@pytest.fixture(scope="session")
def golden_names():
    return ["A_3kbps", "B_3kbps", "A_1p5kbps"]


def pytest_sessionfinish(session, exitstatus):
    """Measured parity figures of the GPU tests -> gpurun_out/parity_report.json (the numbers DESIGN.md §2 quotes)."""
    import json
    try:
        import parity_common as pc
    except Exception:
        return
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out) and pc.MEASURED:
        path = os.path.join(out, "parity_report.json")
        old = json.load(open(path)) if os.path.exists(path) else {}
        old.update(pc.MEASURED)
        with open(path, "w") as f:
            json.dump(old, f, indent=1)
