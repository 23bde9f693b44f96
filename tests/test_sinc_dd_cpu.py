"""CPU test of the product's double-double tap recipe (audio_fir_filter_b200/csrc/sinc_dd.cuh,
compiled here with g++ through tests/harness/sinc_dd_host.cpp): it must reproduce the
50-digit mpmath golden taps BIT FOR BIT, and sit within 1 ulp of the long-double oracle."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, golden


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("h") / "sinc_dd_host")
    src = os.path.join(ROOT, "tests", "harness", "sinc_dd_host.cpp")
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-o", exe, src, "-lm"], capture_output=True,
                       text=True)
    assert r.returncode == 0, r.stderr
    return exe


def dd_taps(exe, fc, M):
    out = subprocess.run([exe, repr(float(fc)), str(int(M))], capture_output=True).stdout
    return np.frombuffer(out, dtype=np.float64)


@pytest.mark.parametrize("name", ["small", "odd", "cfg1", "cfg2"])
def test_dd_recipe_equals_mpmath_golden_bit_for_bit(harness, name):
    g = golden(f"taps_{name}.npz")
    h = dd_taps(harness, g["fc"], g["taps"].size - 1)
    assert np.array_equal(h, g["taps"])
    assert not np.signbit(h[0]) and h[0] == 0.0           # the window's exact zero is +0
    assert np.array_equal(h, h[::-1])


def test_dd_recipe_within_one_ulp_of_the_long_double_oracle(harness, oracle_mod):
    for fs, f, s in [(96000.0, 10.0, 20.0), (192000.0, 15.0, 50.0), (8000.0, 1000.0, 500.0)]:
        want, ld = oracle_mod.build_lowcut(f / fs, s / fs, want_ld=True)
        h = dd_taps(harness, f / fs, want.size - 1)
        err = np.abs(h.astype(np.longdouble) - ld).astype(np.float64)
        assert np.all(err <= np.spacing(np.abs(want)) + 1e-19)
        # the oracle rounds twice (80-bit, then binary64): a fraction of a per cent of
        # its taps are 1 ulp off the correctly rounded value
        assert np.count_nonzero(h != want) <= 0.01 * h.size + 2
