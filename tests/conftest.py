import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    """`-m gpu` tests are skipped (not failed) on a box without a GPU."""
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(autouse=True)
def _product_defaults():
    """Engine options live in module globals (ops.set_precision, engine.FUSED_*): every test starts from, and leaves
    behind, the product configuration, so results do not depend on test order."""
    from geoformer_b200 import engine, ops

    def reset():
        ops.set_precision(linear="tf32", similarity="f16x3", attention="tf32", activations="f16")
        engine.FUSED_MATCHING = True
        engine.FUSED_FINE_LAYER = True
    reset()
    yield
    reset()
