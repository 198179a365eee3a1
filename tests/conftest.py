import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "needs_reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    from oracle import ref_shims
    have_ref = ref_shims.have_reference()
    for item in items:
        if "needs_reference" in item.keywords and not have_ref:
            item.add_marker(pytest.mark.skip(reason="reference tree not present"))


GOLDEN = os.path.join(ROOT, "tests", "golden")
CKPT_DIR = os.path.join(ROOT, "checkpoints", "_ref")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
