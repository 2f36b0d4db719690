import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=False)


@pytest.fixture(scope="session")
def golden():
    return load_golden


@pytest.fixture(scope="session")
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


# Measured parity figures (Jaccard of index sets, tie margins, sigma / Frobenius errors) are collected here and
# printed in the terminal summary, so that they are in the log of a plain `pytest -q` run even when every test passes.
_PARITY_LINES = []


@pytest.fixture
def parity_log():
    def log(line: str):
        _PARITY_LINES.append(line)
    return log


def pytest_terminal_summary(terminalreporter):
    if _PARITY_LINES:
        terminalreporter.write_sep("-", "measured parity against the oracle")
        for line in _PARITY_LINES:
            terminalreporter.write_line(line)
