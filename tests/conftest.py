import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def spec():
    from oracle import satrn
    return satrn.ModelSpec()


@pytest.fixture(scope="session")
def ckpt0(spec):
    from oracle import synth
    return synth.synth_state_dict(spec, 0)


@pytest.fixture(scope="session")
def ckpt1(spec):
    from oracle import synth
    return synth.synth_state_dict(spec, 1)


def load_golden(seed):
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "efficientsatrn_seed%d.npz" % seed))


def load_manager_golden():
    """Reference greedy decode under its DecodingManager (oracle/make_golden.py --manager) + the compiled rule tables."""
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "efficientsatrn_seed0_manager.npz"))
