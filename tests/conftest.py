import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


def load_golden(name):
    """-> dict of torch tensors; keys 'sd.*' are the reference state_dict, 'grad.*' parameter grads."""
    z = np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)
    out = {}
    for k in z.files:
        a = z[k]
        out[k] = torch.from_numpy(a) if a.dtype.kind in 'fiub' and a.ndim > 0 else a
    return out


def state_dict_of(g, tag='sd.'):
    return {k[len(tag):]: v for k, v in g.items() if k.startswith(tag)}


@pytest.fixture(scope='session')
def golden():
    return load_golden


def sse_from_sizes(sizes):
    cs = np.concatenate([[0], np.cumsum(sizes)])
    return torch.tensor([[cs[i], cs[i + 1]] for i in range(len(sizes))], dtype=torch.int64)
