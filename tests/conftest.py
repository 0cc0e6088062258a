import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    torch.set_default_dtype(torch.float64)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


def oracle_flow_from_golden(g):
    """(N,955) flat reference weights -> oracle Flow."""
    from oracle import fthmc_oracle as O
    shapes = [(8, 2, 3, 3), (8,), (8, 8, 3, 3), (8,), (3, 8, 3, 3), (3,)]
    layers = []
    for i, row in enumerate(g["weights"]):
        parts, pos = [], 0
        for shp in shapes:
            n = int(np.prod(shp))
            parts.append(torch.from_numpy(row[pos:pos + n].reshape(shp).copy()))
            pos += n
        assert pos == row.size
        layers.append(O.LayerWeights(w=parts[0::2], b=parts[1::2], mu=i % 2, off=(i // 2) % 4))
    return O.Flow(layers=layers, activation=str(g["activation"]), convention=int(g["convention"]))


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = load_golden(name)
        return cache[name]
    return get


def thousand_inputs(L, n=1000, seed=None):
    """Inputs of the many-trajectory parity runs (regenerated, not stored): see make_golden.thousand_inputs."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN, "make_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.thousand_inputs(L, n) if seed is None else m.thousand_inputs(L, n, seed=seed)
