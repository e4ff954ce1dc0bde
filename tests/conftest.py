"""pytest configuration: the `gpu` marker, import path and shared fixtures."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_NAMES = ["simplevit_tiny", "vit_tiny_cls", "vit_tiny_mean_nonsquare", "vit_tiny_tokens"]
REFERENCE_DIR = os.environ.get("VIT_REFERENCE", "/root/reference")
if not os.path.isdir(os.path.join(REFERENCE_DIR, "vit_pytorch")):
    # GPU box: the unmodified reference installed under baseline/_ref travels with the snapshot
    REFERENCE_DIR = os.path.join(ROOT, "baseline", "_ref")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return torch.load(os.path.join(GOLDEN_DIR, name + ".pt"), weights_only=False)


@pytest.fixture(params=GOLDEN_NAMES)
def golden(request):
    return load_golden(request.param)


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_DIR, "vit_pytorch"))


def import_reference():
    """Import the unmodified reference package (build container only)."""
    import importlib
    sys.dont_write_bytecode = True
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    return importlib.import_module("vit_pytorch")
