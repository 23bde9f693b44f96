"""CPU tests of the host-side logic above the C-ABI (no device work)."""
import pytest

from audio_fir_filter_b200.process import FilterOptions, PcmInfo, plan_blocks, scale_for_peak
from audio_fir_filter_b200.dist import assign_files


def test_filter_options_defaults_match_reference_cli():
    o = FilterOptions()                # main.cp:44-50 defaults
    assert (o.freq, o.slope, o.normalize, o.verbose, o.num_threads) == (15.0, 10.0, False, False, 0)


def test_scale_rule_matches_processfile():
    assert scale_for_peak(0.25, False) == 1.0
    assert scale_for_peak(0.25, True) == 4.0
    assert scale_for_peak(1.0, False) == 1.0
    assert scale_for_peak(2.0, False) == 0.5
    assert scale_for_peak(0.0, True) == 1.0


@pytest.mark.parametrize("frames,world,half", [(1000, 1, 40), (1000, 2, 40), (1003, 3, 400), (5_529_600_000, 8, 76800),
                                               (10, 4, 3), (0, 2, 5), (17, 8, 100)])
def test_plan_blocks_partitions_the_file(frames, world, half):
    blocks = plan_blocks(frames, world, half)
    assert len(blocks) == world
    pos = 0
    for b in blocks:
        assert b.start == pos and b.frames >= 0
        assert b.halo_left == min(half, b.start)
        assert b.halo_right == min(half, frames - b.start - b.frames)
        if b.rank < world - 1 and b.frames and b.start + b.frames < frames:
            assert b.frames % 16 == 0            # seams stay 128-byte aligned in FP64
        pos += b.frames
    assert pos == frames
    sizes = [b.frames for b in blocks if b.frames]
    if sizes:
        assert max(sizes) - min(sizes[:-1] or sizes) <= 16 or len(sizes) < world


def test_pcm_info_sizes():
    i = PcmInfo(frames=10, channels=3, bits=24, big_endian=False, sample_rate=48000.0)
    assert i.frame_bytes == 9 and i.nbytes == 90


def test_assign_files_is_balanced_and_complete():
    sizes = [5, 1, 9, 3, 3, 7, 2, 8]
    per = assign_files(sizes, 3)
    assert sorted(i for lst in per for i in lst) == list(range(len(sizes)))
    loads = [sum(sizes[i] for i in lst) for lst in per]
    assert max(loads) - min(loads) <= max(sizes)
    assert assign_files([4] * 256, 8) == [[i for i in range(256) if i % 8 == r] for r in range(8)]


def test_plan_blocks_random_inputs():
    import random

    rnd = random.Random(11)
    for _ in range(500):
        frames = rnd.choice([0, 1, 15, 16, 17, rnd.randrange(1, 10**6), rnd.randrange(1, 10**10)])
        world = rnd.randrange(1, 17)
        half = rnd.randrange(0, 200_000)
        blocks = plan_blocks(frames, world, half)
        assert len(blocks) == world and sum(b.frames for b in blocks) == frames
        pos = 0
        for b in blocks:
            assert b.start == pos and 0 <= b.halo_left <= min(half, b.start)
            assert b.halo_left == min(half, b.start) and b.halo_right == min(half, frames - b.start - b.frames)
            assert b.start % 16 == 0 or b.frames == 0 or b.start == frames     # seams on 128-byte FP64 boundaries
            pos += b.frames
