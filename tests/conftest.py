import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["default_eth", "odd_small", "deep_blocks", "stress_wide", "tuned_h128_c64", "config5_full"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    from flowchain_iccv2023_b200 import synth
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    s = [int(v) for v in g["spec"]]
    spec = synth.FlowSpec(pred_len=s[0], input_size=s[1], cond_size=s[2], hidden_size=s[3], n_hidden=s[4], n_blocks=s[5])
    params = synth.make_params(spec, int(g["seed"]))
    return spec, params, g


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    return load_golden(request.param)


def update_times(T):
    return sorted({1, T // 2, T - 1})
