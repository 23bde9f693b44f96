"""CPU tests of bench.py's parity CHECKER (bench.ParityWindows): a checker that cannot fail proves
nothing, so it is fed (a) the oracle's own output through a stand-in context -- must pass -- and
(b) the same with one sample pushed just beyond the D3 tolerance, one PCM sample off by one LSB
too many, a seam bug (a block filtered without its halo) -- must fail each time."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402
from audio_fir_filter_b200.process import plan_blocks  # noqa: E402

CFG = dict(name="checker test: stereo 8 kHz 16-bit BE, -n", fs=8000, freq=40.0, slope=50.0, channels=2, bits=16, be=True,
           normalize=True, frames=40_000)


class FakeKernel:
    def __init__(self, taps):
        self._t = taps
        self.half_len = (taps.size - 1) // 2

    def taps(self):
        return self._t


class FakeCtx:
    """Serves windows of a 'parked signal' computed on the CPU (with optional defects)."""

    def __init__(self, y_block):
        self.y = y_block

    def parked_range(self, first, frames, ch):
        return np.ascontiguousarray(self.y[:, first:first + frames])


def make_block(oracle_mod, world, rank, halo=True):
    c = CFG
    taps = oracle_mod.build_lowcut(c["freq"] / c["fs"], c["slope"] / c["fs"])
    H = (taps.size - 1) // 2
    blk = plan_blocks(c["frames"], world, H)[rank]
    pcm = oracle_mod.synth_pcm(bench.SEED, 0, c["frames"], c["channels"], c["bits"], c["be"], c["fs"])
    x = oracle_mod.decode(pcm, c["frames"], c["channels"], c["bits"], c["be"])
    if not halo:                                   # the seam bug: the block is filtered as if it were a whole file
        x = x[:, blk.start:blk.start + blk.frames]
        y = np.stack([oracle_mod.fir_hi(x[k], taps) for k in range(c["channels"])])
    else:
        y = np.stack([oracle_mod.fir_hi(x[k], taps)[blk.start:blk.start + blk.frames] for k in range(c["channels"])])
    return taps, blk, y


def run_checker(oracle_mod, taps, blk, y, pcm_block, scale):
    pw = bench.ParityWindows(CFG, blk, CFG["frames"], bench.SEED, FakeKernel(taps), blk.rank, W=256, n_random=3)
    pw.check_taps(CFG).check_signal(FakeCtx(y)).check_pcm(scale, lambda lo, hi: pcm_block[lo:hi])
    return pw.result


@pytest.mark.parametrize("world,rank", [(1, 0), (2, 0), (2, 1), (3, 1)])
def test_checker_accepts_the_oracles_own_output(oracle_mod, world, rank):
    taps, blk, y = make_block(oracle_mod, world, rank)
    scale = 1.7
    pcm = oracle_mod.encode(y, scale, CFG["bits"], CFG["be"])
    r = run_checker(oracle_mod, taps, blk, y, pcm, scale)
    assert r["ok"] and r["windows"] >= 4 and r["flips"] == 0 and r["worst_d3"] <= 1e-15 and r["taps_max_ulp"] == 0.0


def test_checker_rejects_a_sample_beyond_the_d3_tolerance(oracle_mod):
    taps, blk, y = make_block(oracle_mod, 2, 1)
    y = y.copy()
    y[1, 3] += 1e-10                                # the first window of the block: right at the seam
    pcm = oracle_mod.encode(y, 1.0, CFG["bits"], CFG["be"])
    r = run_checker(oracle_mod, taps, blk, y, pcm, 1.0)
    assert not r["ok"] and r["worst_d3"] > bench.TOL


def test_checker_rejects_a_block_filtered_without_its_halo(oracle_mod):
    """FilterCore.h:57-61,72-76 at a seam: zeros where the neighbour's samples belong."""
    taps, blk, y = make_block(oracle_mod, 2, 1, halo=False)
    pcm = oracle_mod.encode(y, 1.0, CFG["bits"], CFG["be"])
    r = run_checker(oracle_mod, taps, blk, y, pcm, 1.0)
    assert not r["ok"] and r["worst_d3"] > 1e-6     # the first window after the seam is plainly wrong


def test_checker_rejects_wrong_pcm_and_wrong_scale(oracle_mod):
    taps, blk, y = make_block(oracle_mod, 1, 0)
    pcm = oracle_mod.encode(y, 1.0, CFG["bits"], CFG["be"]).copy()
    ok = run_checker(oracle_mod, taps, blk, y, pcm, 1.0)
    assert ok["ok"]
    bad = pcm.copy()
    bad[1] ^= 0x02                                  # frame 0, channel 0, low byte (big-endian): 2 LSB off
    r = run_checker(oracle_mod, taps, blk, y, bad, 1.0)
    assert not r["ok"] and r["max_flip_lsb"] == 2
    # the PCM of another scale (a peak that was not all-reduced, ProcessFile.cp:92-101)
    r = run_checker(oracle_mod, taps, blk, y, pcm, 1.001)
    assert not r["ok"] and r["flips"] > 4


def test_checker_rejects_taps_that_are_more_than_one_ulp_off(oracle_mod):
    taps, blk, y = make_block(oracle_mod, 1, 0)
    t2 = taps.copy()
    t2[10] = np.nextafter(np.nextafter(t2[10], 1.0), 1.0)
    t2[-11] = t2[10]                                # keep the symmetry: only the ulp test can object
    pw = bench.ParityWindows(CFG, blk, CFG["frames"], bench.SEED, FakeKernel(t2), 0, W=256, n_random=1)
    pw.check_taps(CFG)
    assert not pw.result["ok"] and pw.result["taps_max_ulp"] == 2.0


def test_both_arms_build_the_same_config_object():
    for cid, cfg in bench.CONFIGS.items():
        for world in (1, 2, 8):
            a = bench.config_dict(cfg, world, "block")
            assert a["taps"] == bench.kernel_order(cfg["slope"] / cfg["fs"]) + 1 and a["frames_per_gpu"] == cfg["frames"]
            assert ("N=%d" % world in a["workload"]) == (world > 1)
    s = bench.config_dict(bench.CONFIGS[3], 8, "block", strong=True)
    assert s["frames_per_gpu"] == bench.CONFIGS[3]["frames"] // 8 and "split into 8 sample blocks" in s["workload"]
