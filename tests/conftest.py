import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def gold_dir():
    return GOLD


@pytest.fixture(scope="session")
def coin_label_emb():
    import torch
    e = torch.load(os.path.join(GOLD, "clip_step_emb_coin.pt"))
    return e / e.norm(dim=1, keepdim=True)      # GPU semantics of check_device_norm (vit.py:435-440)
