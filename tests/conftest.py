"""pytest configuration: the ``gpu`` marker and shared fixtures.

``-m "not gpu"`` runs here (no GPU): oracle vs golden vectors, host logic, C-ABI symbol table.
``-m gpu`` runs on a B200: the parity tests proper, every one calling through the C-ABI library.
"""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason='no CUDA device')
    for it in items:
        if 'gpu' in it.keywords:
            it.add_marker(skip)


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), map_location='cpu', weights_only=False)


@pytest.fixture(scope='session')
def golden_functions():
    return load_golden('functions.pt')


@pytest.fixture(scope='session')
def golden_layers():
    return load_golden('layers.pt')


@pytest.fixture(scope='session')
def golden_nets():
    return load_golden('nets.pt')
