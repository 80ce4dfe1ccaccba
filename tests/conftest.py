import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a real B200 (run with -m gpu on the GPU box)')


@pytest.fixture(scope='session')
def built_lib():
    """The in-tree shared library (built on demand; nvcc cross-compiles without a GPU)."""
    from fastdepth_b200 import build
    from oracle import build_oracle
    build_oracle.build()                      # the plain-C oracle primitives ride along (gcc, test infrastructure)
    return build.build()


def rel_err(got, want):
    """max |got-want| / max(|want|, mean|want|): element-wise relative error with a floor at the
    tensor's mean magnitude, so exact zeros after ReLU do not blow the ratio up."""
    import torch
    got = got.double(); want = want.double()
    denom = torch.maximum(want.abs(), want.abs().mean())
    return ((got - want).abs() / denom).max().item()
