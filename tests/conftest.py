"""pytest configuration: markers and import paths.

`-m "not gpu"` runs on the CPU-only build container (oracle vs golden vectors,
host logic, C-ABI symbol export); `-m gpu` are the parity tests proper and need
a B200 (run through `gpurun`).
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run under gpurun")


@pytest.fixture(scope="session")
def oracle():
    from oracle import cqs_oracle

    return cqs_oracle
