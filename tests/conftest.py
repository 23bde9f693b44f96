import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def golden(name: str):
    return np.load(os.path.join(GOLDEN, name))


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle

    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def ctx():
    """One C-ABI context on cuda:0 for the GPU tests (no fallback: fails without a B200)."""
    from audio_fir_filter_b200 import capi

    c = capi.Context(0)
    yield c
    c.close()


# The five BASELINE.json configs: (fs, freq, slope, channels, bits, big_endian, normalize, frames)
CONFIGS = {
    1: dict(fs=48000, freq=20.0, slope=20.0, channels=2, bits=24, be=False, normalize=False, frames=2_880_000, taps=9601),
    2: dict(fs=44100, freq=30.0, slope=10.0, channels=2, bits=16, be=True, normalize=True, frames=26_460_000, taps=17641),
    3: dict(fs=96000, freq=10.0, slope=2.0, channels=8, bits=24, be=False, normalize=False, frames=345_600_000, taps=192001),
    4: dict(fs=48000, freq=20.0, slope=20.0, channels=2, bits=24, be=False, normalize=False, frames=14_400_000, taps=9601),
    5: dict(fs=192000, freq=15.0, slope=5.0, channels=16, bits=32, be=False, normalize=True, frames=5_529_600_000, taps=153601),
}
