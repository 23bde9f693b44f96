"""GPU tests at BASELINE.json's full sizes.  The oracle cannot filter 10^12..10^15 MACs, so
parity at these sizes goes through size-independent properties:

  * the input is the counter-based synthetic PCM, generated ON the device and reproducible
    on the host for any window -- sampled windows of the parked FP64 signal (both file ends,
    random interior, every block seam) are checked against the long-double oracle to 1e-12;
  * the encoded PCM of those windows is checked against the oracle's encode (counted flips);
  * the fused peak equals the stand-alone peak kernel's;
  * DC rejection: the 0.05 FS offset of the input is gone from every interior window;
  * -n brings the loudest sample to full scale.
"""
import numpy as np
import pytest

from conftest import CONFIGS
from test_gpu_parity import TOL, lsb_flips, pcm_to_int

pytestmark = pytest.mark.gpu
SEED = 0xF1F1F1


def check_windows(ctx, oracle_mod, k, first_frame, frames, total_frames, ch, bits, be, fs, starts, W, out_dev=None,
                  scale=1.0):
    """Windows [s, s+W) of the block that starts at absolute frame `first_frame` (block-local
    s), against the oracle run on the absolute-position synthetic PCM."""
    taps = k.taps()
    H = k.half_len
    fb = ch * bits // 8
    worst = 0.0
    flips = 0
    for s in starts:
        a = first_frame + s                                   # absolute frame of the window
        lo, hi = max(0, a - H), min(total_frames, a + W + H)  # clipped only at the true file ends
        seg = oracle_mod.synth_pcm(SEED, lo, hi - lo, ch, bits, be, fs)
        x = oracle_mod.decode(seg, hi - lo, ch, bits, be)
        got = ctx.parked_range(s, W, ch)
        want = np.empty_like(got)
        for c in range(ch):
            want[c] = oracle_mod.fir_hi(x[c], taps, a - lo, a - lo + W)[a - lo:a - lo + W]
            sc = oracle_mod.fir_abs_scale(x[c], taps, a - lo, a - lo + W)[a - lo:a - lo + W]
            err = np.abs(got[c] - want[c])
            assert np.all(err <= TOL * sc), (s, c, float((err / sc).max()))
            worst = max(worst, float((err / sc).max()))
        if out_dev is not None:
            pcm_got = out_dev[s * fb:(s + W) * fb].cpu().numpy()
            pcm_want = oracle_mod.encode(want, scale, bits, be)
            n, mx = lsb_flips(pcm_got, pcm_want, bits, be)
            assert mx <= 1 and n <= 2, (s, n, mx)
            flips += n
    return worst, flips


def run_config(ctx, oracle_mod, cfg_id, frames=None, W=1024, n_random=10):
    import torch

    c = CONFIGS[cfg_id]
    frames = frames or c["frames"]
    ch, bits, be, fs = c["channels"], c["bits"], c["be"], c["fs"]
    fb = ch * bits // 8
    d_in = torch.empty(frames * fb, dtype=torch.uint8, device="cuda:0")
    ctx.synth_pcm_dev(SEED, 0, frames, ch, bits, be, fs, 1.0, d_in)
    k = ctx.build_kernel(c["freq"] / fs, c["slope"] / fs)
    assert k.num_taps == c["taps"]
    ctx.apply_dev(k, d_in, frames, ch, bits, be)
    pk = ctx.peak()
    t = ctx.last_timing()
    assert 0.0 < pk < 1.0
    assert ctx.peak_recompute() == pk
    scale = 1.0 / pk if c["normalize"] else 1.0
    d_out = torch.empty_like(d_in)
    ctx.encode_dev(scale, d_out)
    ctx.synchronize()
    rng = np.random.default_rng(cfg_id)
    H = k.half_len
    starts = [0, frames - W] + [int(s) for s in rng.integers(H, frames - H - W, n_random)]
    worst, flips = check_windows(ctx, oracle_mod, k, 0, frames, frames, ch, bits, be, fs, starts, W, d_out, scale)
    # DC rejection: one second (whole periods of the 5 Hz and 1 kHz components) from the middle
    # of the file averages to ~0 although the input carries a 0.05 FS offset
    sec = ctx.parked_range(frames // 2, fs, ch)
    assert np.all(np.abs(sec.mean(axis=1)) < 1e-3), sec.mean(axis=1)
    flop = 2.0 * ch * (c["taps"] * frames - H * (H + 1))
    print(f"config {cfg_id}: {frames} frames x {ch} ch, {c['taps']} taps: fir {t['fir_ms']:.1f} ms "
          f"({flop / t['fir_ms'] / 1e9:.1f} TFLOP/s), worst window error {worst:.2e} of the D3 scale, "
          f"{flips} 1-LSB flips in {len(starts) * W * ch} checked samples")
    if c["normalize"]:
        # -n: some sample of the file sits at full scale.  Scan the encoded PCM on the device.
        if d_out.numel() <= (1 << 28):
            q = pcm_to_int(d_out.cpu().numpy(), bits, be)
            assert max(int(q.max()), -int(q.min())) >= (1 << (bits - 1)) - 1
    k.free()
    del d_in, d_out
    torch.cuda.empty_cache()


def test_config1_full(ctx, oracle_mod):
    run_config(ctx, oracle_mod, 1)


def test_config2_full(ctx, oracle_mod):
    run_config(ctx, oracle_mod, 2)


def test_config4_one_file_full(ctx, oracle_mod):
    run_config(ctx, oracle_mod, 4)


def test_config3_full_one_hour_eight_channels(ctx, oracle_mod):
    """1 h x 8 ch x 96 kHz, 192 001 taps: 1.06e15 FLOP, ~30 s of one B200; streams through the
    bounded FP64 input scratch in ~11 chunks and parks 22 GB."""
    run_config(ctx, oracle_mod, 3, W=512, n_random=12)


def test_config5_slice_as_two_sample_blocks(ctx, oracle_mod):
    """Config 5's shape (16 ch, 192 kHz, 32-bit, 153 601 taps, -n) on a 40 s slice of the 8 h
    file, split into two sample blocks with (taps-1) halo on two contexts (standing in for two
    ranks): windows on both sides of the seam, one common scale from the max of the two peaks."""
    import torch

    from audio_fir_filter_b200 import Context, plan_blocks, scale_for_peak

    c = CONFIGS[5]
    ch, bits, be, fs = c["channels"], c["bits"], c["be"], c["fs"]
    fb = ch * bits // 8
    total = 7_680_000                          # 40 s
    base = 1_000_000_000                       # absolute position inside the 8 h file (no true edge)
    file_frames = c["frames"]
    W = 512
    ctxs = [ctx, Context(0)]
    ks = [cx.build_kernel(c["freq"] / fs, c["slope"] / fs) for cx in ctxs]
    H = ks[0].half_len
    assert ks[0].num_taps == c["taps"]
    blocks = plan_blocks(total, 2, H)
    peaks, outs = [], []
    for b, cx, k in zip(blocks, ctxs, ks):
        # inside the long file both neighbours exist: full halo either side of the 40 s slice too
        hl, hr = H, H
        d_in = torch.empty((hl + b.frames + hr) * fb, dtype=torch.uint8, device="cuda:0")
        cx.synth_pcm_dev(SEED, base + b.start - hl, hl + b.frames + hr, ch, bits, be, fs, 1.0, d_in)
        cx.apply_dev(k, d_in, b.frames, ch, bits, be, hl, hr)
        peaks.append(cx.peak())
        del d_in
    scale = scale_for_peak(max(peaks), True)
    for b, cx, k in zip(blocks, ctxs, ks):
        d_out = torch.empty(b.frames * fb, dtype=torch.uint8, device="cuda:0")
        cx.encode_dev(scale, d_out)
        cx.synchronize()
        starts = [0, b.frames - W, b.frames // 2] + [int(v) for v in np.random.default_rng(b.rank).integers(0, b.frames - W, 5)]
        worst, flips = check_windows(cx, oracle_mod, k, base + b.start, b.frames, file_frames, ch, bits, be, fs,
                                     starts, W, d_out, scale)
        print(f"config 5 block {b.rank}: worst {worst:.2e}, {flips} flips, fir {cx.last_timing()['fir_ms']:.0f} ms")
        del d_out
    for k in ks:
        k.free()
    ctxs[1].close()
    torch.cuda.empty_cache()
