import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
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


def reference_path():
    """Where an importable copy of the reference (`srl`) is: /root/reference in the build container; baseline/_ref (the offline
    install made by __graft_entry__.build(), git-ignored, shipped to the GPU box with the snapshot) elsewhere; None if neither."""
    for p in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if os.path.isfile(os.path.join(p, "srl", "__init__.py")):
            return p
    return None


@pytest.fixture(scope="session")
def srl_mod():
    """(dqn, rainbow) modules of the reference, or skip."""
    ref = reference_path()
    if ref is None:
        pytest.skip("reference not present")
    if ref not in sys.path:
        sys.path.insert(0, ref)
    import srl  # noqa: F401
    from srl.algorithms import dqn, rainbow

    return dqn, rainbow
