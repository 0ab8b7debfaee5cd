import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


_PARITY_LOG = {}


@pytest.fixture(scope="session")
def parity_log():
    """Measured parity errors per case, written to gpurun_out/parity_log.json (or $LSTEP_PARITY_LOG) at session end;
    the committed copy is profiles/r02_parity_errors.json."""
    return _PARITY_LOG


def pytest_sessionfinish(session, exitstatus):
    if not _PARITY_LOG:
        return
    import json
    path = os.environ.get("LSTEP_PARITY_LOG") or os.path.join(ROOT, "gpurun_out", "parity_log.json")
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as f:
            json.dump(_PARITY_LOG, f, indent=1, sort_keys=True)
    except OSError:
        pass
