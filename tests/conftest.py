import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    out = {}
    for name in ("golden_v1.npz", "golden_v2.npz", "golden_v3.npz", "golden_v4.npz"):     # oracle/make_golden{,_eval,_chain,_c1}.py
        with np.load(os.path.join(ROOT, "tests", "golden", name)) as z:
            out.update({k: z[k] for k in z.files})
    return out


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from downsampled_diffusion_b200 import _lib
    _lib.lib()          # fail loudly if the extension is missing
    return torch.device("cuda:0")
