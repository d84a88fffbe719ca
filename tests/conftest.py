import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def golden_names():
    """model fixtures (tests/golden/make_golden.py); roc_cases.npz belongs to tests/test_roc_golden.py, odin_*.npz to
    tests/test_gpu_odin.py, wim_*.npz to tests/test_gpu_wim.py, cat_*.npz to tests/test_gpu_categorical.py"""
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith('.npz') and not f.startswith(('roc_', 'odin_', 'wim_', 'cat_', 'sig_', 'ycoded_', 'misclass_', 'multistep_', 'full_')))


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN


@pytest.fixture(scope='session')
def pkg():
    """the product package (joint-vae_b200/), with libjvae_sm100.so built/loaded"""
    import __graft_entry__ as g
    return g.build()
