import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.dirname(os.path.abspath(__file__)) not in sys.path:      # test helpers (verify_lib.py) import by module name
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def hp():
    from video_gcp_b200 import hparams
    return hparams.build_hparams(hparams.gcp_tree_25room_config(batch_size=1, attach_cost_mdl=True))


@pytest.fixture(scope="session")
def sd(hp):
    """Seeded synthetic weights (seed 1 = the seed the golden fixtures were generated with)."""
    from video_gcp_b200.synthetic import synthetic_state_dict
    return synthetic_state_dict(hp, 1)
