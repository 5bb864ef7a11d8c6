import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def small_pair():
    """2 x ~1.2 Mbp, 3 contigs, 5 % diverged with SV breaks -- the oracle finishes in seconds"""
    from fastga_b200 import formats, synth
    A, B = synth.make_pair(11, 1_200_000, 3, 0.05, sv_every=60000)
    return formats.genome_from_arrays(A), formats.genome_from_arrays(B)
