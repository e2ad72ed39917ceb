"""pytest configuration: registers the `gpu` marker and wires the import roots.

Import roots (mirroring the reference's two spellings):
  <repo>/gcn-max-cut_b200          -> `from python.Training.TrainingNeural import ...`, `import gmc_b200`
  <repo>/gcn-max-cut_b200/python   -> `from Training.TrainingNeural import ...`, `from commons import ...`
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "gcn-max-cut_b200")
for p in (ROOT, PKG, os.path.join(PKG, "python")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
