import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN_DIR = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    import numpy as np
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    meta = [str(m) for m in g.pop("meta")]
    g["cfg_name"], g["antialias"], g["num_samples"], g["batch"], g["kind"], g["seed"] = (
        meta[0], bool(int(meta[1])), int(meta[2]), int(meta[3]), meta[4], int(meta[5]))
    return g


GOLDEN_CASES = ["tiny", "tiny_aa", "tiny_ragged", "base_1s", "base_aa_1s", "debug_1s",
                "debug_causal_1s", "debug_nodil_1s", "config9_base_1s", "default_half_s"]


@pytest.fixture(scope="session")
def golden_loader():
    return load_golden
