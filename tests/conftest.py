import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def oracle_model():
    """(layers, save, unfused state_dict, fused dict) of the seed-0 calibrated synthetic Rep-YOLO (CPU oracle)."""
    from oracle import repyolo_oracle as O
    return O.make_model(seed=0, mode='calibrated')
