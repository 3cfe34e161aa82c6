import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with `-m gpu`)')


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not errored) on a machine without a CUDA device."""
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason='needs a CUDA device')
    for item in items:
        if 'gpu' in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope='session')
def native_lib():
    """The C-ABI library; (re)built in-tree when stale.  Loading it does not need a GPU."""
    from demuxalot_b200 import _native, build
    build.build()
    return _native.load()
