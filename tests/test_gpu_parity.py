"""GPU parity tests: the CUDA path, called through the C-ABI, against the oracle.

Bars (BASELINE.json north_star; decisions D1-D5 in DESIGN.md):
  * FP64 filtered signal: |y_gpu - y_oracle| <= 1e-12 * sum_k |h_k x_{n-H+k}|      (D3)
  * taps: within 1 ulp of the oracle's long-double taps
  * decode: exact;  encode / whole path: PCM bit-exact except a COUNTED number of
    +-1-LSB flips at rounding boundaries (asserted to be a tiny fraction)
  * tiling / chunking / variant / sharding invariance: bit-exact
"""
import numpy as np
import pytest

from conftest import CONFIGS, golden

pytestmark = pytest.mark.gpu

TOL = 1e-12
TAP_ULPS = 1.0   # device taps (double-double, one rounding) vs the oracle's long-double taps


def pcm_to_int(pcm, bits, be):
    nb = bits // 8
    b = pcm.reshape(-1, nb).astype(np.int64)
    if be:
        b = b[:, ::-1]
    v = np.zeros(b.shape[0], dtype=np.int64)
    for k in range(nb):
        v |= b[:, k] << (8 * k)
    sign = 1 << (bits - 1)
    return (v ^ sign) - sign


def lsb_flips(a, b, bits, be):
    """(#samples that differ, max |difference| in LSB)."""
    d = np.abs(pcm_to_int(a, bits, be) - pcm_to_int(b, bits, be))
    return int(np.count_nonzero(d)), int(d.max(initial=0))


# ------------------------------------------------------------------ taps ----------

@pytest.mark.parametrize("cfg", [1, 2, 3, 5])
def test_build_kernel_matches_oracle_taps(ctx, oracle_mod, cfg):
    c = CONFIGS[cfg]
    fc, bw = c["freq"] / c["fs"], c["slope"] / c["fs"]
    k = ctx.build_kernel(fc, bw)
    assert k.num_taps == c["taps"] and k.half_len == (c["taps"] - 1) // 2
    got = k.taps()
    want, ld = oracle_mod.build_lowcut(fc, bw, want_ld=True)
    # the device carries every tap in double-double and rounds once; the oracle rounds
    # twice (80-bit, then binary64): they differ by at most 1 ulp, on well under 1 % of the taps
    err = np.abs(got.astype(np.longdouble) - ld).astype(np.float64)
    assert np.all(err <= TAP_ULPS * np.spacing(np.abs(want))), float((err / np.spacing(np.abs(want))).max())
    ndiff = int(np.count_nonzero(got != want))
    print(f"config {cfg}: {ndiff} of {got.size} taps differ from the oracle's (by 1 ulp)")
    assert ndiff <= 0.01 * got.size
    assert np.array_equal(got, got[::-1])                       # symmetric bit for bit
    assert abs(float(got.astype(np.longdouble).sum())) < 1e-15   # DC removed
    assert got[0] == 0.0 and not np.signbit(got[0])
    k.free()


@pytest.mark.parametrize("name", ["small", "odd", "cfg1", "cfg2"])
def test_build_kernel_equals_mpmath_golden_bit_for_bit(ctx, name):
    """The 50-digit mpmath evaluation rounded once to binary64 is the correctly rounded tap;
    the device's double-double recipe reproduces it exactly."""
    g = golden(f"taps_{name}.npz")
    k = ctx.build_kernel(float(g["fc"]), float(g["bw"]))
    got = k.taps()
    assert np.array_equal(got, g["taps"])
    k.free()


def test_build_kernel_device_equals_host_compile_of_the_same_recipe(ctx, tmp_path):
    """sinc_dd.cuh compiled with g++ (tests/harness) and with nvcc give the same taps for the
    long kernels of configs 3 and 5 (no golden is stored for those: 1.5 MB each)."""
    import os
    import subprocess

    from conftest import ROOT

    exe = str(tmp_path / "sinc_dd_host")
    r = subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-o", exe,
                        os.path.join(ROOT, "tests", "harness", "sinc_dd_host.cpp"), "-lm"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    for cfg in (3, 5):
        c = CONFIGS[cfg]
        fc, bw = c["freq"] / c["fs"], c["slope"] / c["fs"]
        k = ctx.build_kernel(fc, bw)
        got = k.taps()
        out = subprocess.run([exe, repr(fc), str(got.size - 1)], capture_output=True).stdout
        assert np.array_equal(got, np.frombuffer(out, dtype=np.float64))
        k.free()


def test_build_kernel_rejects_bad_arguments(ctx):
    from audio_fir_filter_b200 import capi

    for fc, bw in [(0.01, 0.0), (0.01, -1.0), (0.0, 0.01), (0.5, 0.01), (0.01, 1e-12)]:
        with pytest.raises(capi.FirGpuError) as e:
            ctx.build_kernel(fc, bw)
        assert e.value.code == capi.ERR_INVALID


# ------------------------------------------------------------------- FIR ----------

def check_fir(ctx, oracle_mod, taps, x):
    k = ctx.kernel_from_taps(taps)
    y = ctx.filter_f64(k, x)
    k.free()
    x2 = np.atleast_2d(x)
    worst = 0.0
    for c in range(x2.shape[0]):
        want = oracle_mod.fir_hi(x2[c], taps)
        scale = oracle_mod.fir_abs_scale(x2[c], taps)
        err = np.abs(y[c] - want)
        assert np.all(err <= TOL * scale + 1e-300), float((err / np.maximum(scale, 1e-300)).max())
        worst = max(worst, float((err / np.maximum(scale, 1e-300)).max()))
    return y, worst


@pytest.mark.parametrize("N", [1, 2, 15, 16, 17, 480, 961, 962, 4095, 4096, 4097, 10000, 70001])
def test_fir_f64_any_length_including_shorter_than_the_kernel(ctx, oracle_mod, N):
    taps = oracle_mod.build_lowcut(20.0 / 48000, 200.0 / 48000)      # 961 taps
    x = np.random.default_rng(N).uniform(-1, 1, N)
    check_fir(ctx, oracle_mod, taps, x)


def test_fir_f64_multichannel_long_kernel(ctx, oracle_mod):
    taps = oracle_mod.build_lowcut(20.0 / 48000, 20.0 / 48000)       # config 1 kernel, 9601 taps
    x = np.random.default_rng(3).uniform(-1, 1, (3, 40000))
    _, worst = check_fir(ctx, oracle_mod, taps, x)
    assert worst < 1e-13


def test_fir_f64_against_reference_filtercore_golden(ctx, oracle_mod):
    for name in ("body", "maketest", "short"):
        g = golden(f"filtercore_{name}.npz")
        k = ctx.kernel_from_taps(g["taps"])
        y = ctx.filter_f64(k, g["x"].astype(np.float64))[0]
        k.free()
        ulp = np.spacing(np.abs(g["y_full"])).astype(np.float64)
        # the reference narrows to float32 on store (FilterCore.h:59,67,74)
        assert np.all(np.abs(y - g["y_full"]) <= 0.5 * ulp + 1e-12)
        flips = np.count_nonzero(y.astype(np.float32) != g["y_full"])
        assert flips <= 0.002 * y.size, flips


def test_fir_impulse_returns_the_taps_exactly(ctx, oracle_mod):
    taps = oracle_mod.build_lowcut(30.0 / 44100, 100.0 / 44100)
    M = taps.size - 1
    H = M // 2
    N = 3 * M
    x = np.zeros(N)
    p = N // 2 + 3
    x[p] = 1.0
    k = ctx.kernel_from_taps(taps)
    y = ctx.filter_f64(k, x)[0]
    assert np.array_equal(y[p - H:p + H + 1], taps[::-1])
    assert not y[:p - H].any() and not y[p + H + 1:].any()
    # constant -> 0 in the steady state, gain ~ 1 at Nyquist
    y = ctx.filter_f64(k, np.full(N, 0.5))[0]
    assert np.max(np.abs(y[H:N - H])) < 1e-15
    y = ctx.filter_f64(k, (-1.0) ** np.arange(N))[0]
    assert np.max(np.abs(np.abs(y[H:N - H]) - 1.0)) < 1e-3
    k.free()


def test_fir_is_invariant_to_variant_and_chunking(ctx, oracle_mod):
    """Within a kernel family (DFMA: one ascending FMA chain per output; DMMA: ascending
    8-tap Toeplitz steps grouped by n mod 8) the bits do not depend on the CTA shape, the
    tiles per warp, the tap-tile size or the pipeline depth; across the two families the
    results agree to the parity tolerance."""
    from audio_fir_filter_b200 import capi

    taps = oracle_mod.build_lowcut(20.0 / 48000, 60.0 / 48000)       # 3201 taps
    x = np.random.default_rng(11).uniform(-1, 1, (2, 50000))
    scale = np.stack([oracle_mod.fir_abs_scale(x[c], taps) for c in range(2)])
    want = np.stack([oracle_mod.fir_hi(x[c], taps) for c in range(2)])
    k = ctx.kernel_from_taps(taps)
    base = {}
    try:
        for v, name in enumerate(capi.variant_names()):
            ctx.set_variant(v)
            y = ctx.filter_f64(k, x)
            fam = name.split("_")[0]
            assert fam in ("dfma", "dmma")
            assert np.array_equal(y, base.setdefault(fam, y)), name
            assert np.all(np.abs(y - want) <= TOL * scale), name
        assert len(base) == 2
    finally:
        ctx.set_variant(0)
        k.free()


# ------------------------------------------------------------------- PCM ----------

@pytest.mark.parametrize("bits", [16, 24, 32])
@pytest.mark.parametrize("be", [False, True])
@pytest.mark.parametrize("channels", [1, 2, 3, 8, 16])
def test_pcm_path_all_formats(ctx, oracle_mod, bits, be, channels):
    fs, freq, slope = 8000, 40.0, 50.0                                # 641 taps
    frames = 5000 + 7 * channels
    pcm = oracle_mod.synth_pcm(bits * 100 + channels, 0, frames, channels, bits, be, fs)
    k = ctx.build_kernel(freq / fs, slope / fs)
    ctx.apply(k, pcm, frames, channels, bits, be)
    # decode + FIR: the parked FP64 signal against the oracle on the oracle's decode
    y = ctx.parked(frames, channels)
    taps = k.taps()
    x = oracle_mod.decode(pcm, frames, channels, bits, be)
    for c in range(channels):
        want = oracle_mod.fir_hi(x[c], taps)
        scale = oracle_mod.fir_abs_scale(x[c], taps)
        assert np.all(np.abs(y[c] - want) <= TOL * scale + 1e-300)
    # peak: fused epilogue value == stand-alone kernel == max|y| of what is parked
    pk = ctx.peak()
    assert pk == float(np.abs(y).max())
    assert ctx.peak_recompute() == pk
    # encode: bit-exact against the oracle's encode of the SAME parked signal
    for scale_ in (1.0, 1.0 / pk, 3.0):                               # 3.0 clips
        out = np.empty_like(pcm)
        ctx.encode(scale_, out)
        want = oracle_mod.encode(y, scale_, bits, be)
        n, mx = lsb_flips(out, want, bits, be)
        assert mx <= 1 and n <= max(2, 1e-4 * out.size), (n, mx)
    k.free()


def test_decode_is_exact_identity_kernel(ctx, oracle_mod):
    """A 1-tap kernel h = [1] turns apply+encode into decode -> encode: bytes in == bytes out."""
    for bits, be, ch in [(16, True, 2), (24, False, 5), (32, False, 16), (24, True, 1)]:
        frames = 33333
        pcm = np.random.default_rng(bits + ch).integers(0, 256, frames * ch * bits // 8, dtype=np.uint8)
        k = ctx.kernel_from_taps(np.array([1.0]))
        ctx.apply(k, pcm, frames, ch, bits, be)
        y = ctx.parked(frames, ch)
        assert np.array_equal(y, oracle_mod.decode(pcm, frames, ch, bits, be))
        out = np.zeros_like(pcm)
        ctx.encode(1.0, out)
        assert np.array_equal(out, pcm)
        k.free()


@pytest.mark.parametrize("cfg,frames", [(1, 300_000), (2, 400_000)])
def test_whole_path_reduced_configs(ctx, oracle_mod, cfg, frames):
    """Configs 1 and 2 (real kernels, real formats, -n on config 2) on a shortened file,
    whole path against oracle_process: y within 1e-12, PCM bit-exact modulo counted flips."""
    from audio_fir_filter_b200 import FilterOptions, PcmInfo, process_pcm

    c = CONFIGS[cfg]
    pcm = oracle_mod.synth_pcm(0xF1F1F1, 0, frames, c["channels"], c["bits"], c["be"], c["fs"])
    info = PcmInfo(frames, c["channels"], c["bits"], c["be"], float(c["fs"]))
    opts = FilterOptions(freq=c["freq"], slope=c["slope"], normalize=c["normalize"])
    out = np.empty_like(pcm)
    r = process_pcm(ctx, pcm, info, opts, out)
    assert r["taps"] == c["taps"]
    want = oracle_mod.process(pcm, frames, c["channels"], c["bits"], c["be"], c["freq"] / c["fs"],
                              c["slope"] / c["fs"], c["normalize"])
    y = ctx.parked(frames, c["channels"])
    # the stated bar (decision D3): per sample, relative to sum_k |h_k * x_k| -- the oracle's taps
    # (<= 1 ulp from the device's) and the decoded input give that scale
    taps = oracle_mod.build_lowcut(c["freq"] / c["fs"], c["slope"] / c["fs"])
    x = oracle_mod.decode(pcm, frames, c["channels"], c["bits"], c["be"])
    for ch in range(c["channels"]):
        d3 = oracle_mod.fir_abs_scale(x[ch], taps)
        assert np.all(np.abs(y[ch] - want["y"][ch]) <= TOL * d3), float((np.abs(y[ch] - want["y"][ch]) / d3).max())
    assert abs(r["peak"] - want["peak"]) <= 1e-12 * want["peak"]
    assert abs(r["scale"] - want["scale"]) <= 1e-12 * want["scale"]
    n, mx = lsb_flips(out, want["pcm"], c["bits"], c["be"])
    print(f"config {cfg}: {n} of {out.size // (c['bits'] // 8)} samples differ by 1 LSB")
    assert mx <= 1 and n <= 1e-5 * out.size + 2, (n, mx)
    if c["normalize"]:
        v = pcm_to_int(out, c["bits"], c["be"])
        assert max(v.max(), -v.min()) >= (1 << (c["bits"] - 1)) - 1


def test_auto_normalise_fires_only_above_full_scale(ctx, oracle_mod):
    from audio_fir_filter_b200 import FilterOptions, PcmInfo, process_pcm

    fs, frames, ch, bits = 8000, 20000, 2, 24
    opts = FilterOptions(freq=40.0, slope=50.0)
    info = PcmInfo(frames, ch, bits, False, float(fs))
    quiet = oracle_mod.synth_pcm(5, 0, frames, ch, bits, False, fs, gain=1.0)
    out = np.empty_like(quiet)
    r = process_pcm(ctx, quiet, info, opts, out)
    assert r["peak"] <= 1.0 and r["scale"] == 1.0
    # a full-scale square wave overshoots after the high-pass (Gibbs): peak > 1 -> scaled down
    sq = np.where((np.arange(frames) // 50) % 2 == 0, (1 << 23) - 1, -(1 << 23)).astype(np.int32)
    b = np.stack([(sq >> s) & 0xFF for s in (0, 8, 16)], axis=1).astype(np.uint8)
    loud = np.repeat(b[:, None, :], ch, axis=1).reshape(-1)
    r = process_pcm(ctx, loud, info, opts, out)
    assert r["peak"] > 1.0 and r["scale"] == 1.0 / r["peak"]
    want = oracle_mod.process(loud, frames, ch, bits, False, 40.0 / fs, 50.0 / fs, False)
    n, mx = lsb_flips(out, want["pcm"], bits, False)
    assert mx <= 1 and n <= 4


# ------------------------------------------------------- sharding / chunking ------

def test_sample_block_sharding_is_bit_exact(ctx, oracle_mod):
    """One long file as 1, 2, 3 and 8 sample blocks with (taps-1) halo (several contexts
    on this one GPU standing in for the ranks): same peak, same PCM, bit for bit."""
    from audio_fir_filter_b200 import Context, FilterOptions, PcmInfo, process_pcm, process_pcm_sharded

    c = CONFIGS[5]
    fs, frames, ch, bits = 8000, 60_003, 4, 32
    pcm = oracle_mod.synth_pcm(9, 0, frames, ch, bits, False, fs)
    info = PcmInfo(frames, ch, bits, False, float(fs))
    opts = FilterOptions(freq=c["freq"] * fs / c["fs"], slope=c["slope"] * fs / c["fs"] * 40, normalize=True)
    whole = np.empty_like(pcm)
    r0 = process_pcm(ctx, pcm, info, opts, whole)
    y0 = ctx.parked(frames, ch)
    for world in (2, 3, 8):
        ctxs = [Context(0) for _ in range(world)]
        out, r = process_pcm_sharded(ctxs, pcm, info, opts)
        assert r["peak"] == r0["peak"] and r["scale"] == r0["scale"]
        assert np.array_equal(out, whole)
        for b in r["blocks"]:
            if b.frames:
                assert np.array_equal(ctxs[b.rank].parked(b.frames, ch), y0[:, b.start:b.start + b.frames])
        for cx in ctxs:
            cx.close()


def test_chunked_streaming_is_bit_exact(ctx, oracle_mod):
    """Files larger than the decoded-input scratch stream through it chunk by chunk."""
    fs, frames, ch, bits = 8000, 300_000, 2, 16
    pcm = oracle_mod.synth_pcm(21, 0, frames, ch, bits, True, fs)
    k = ctx.build_kernel(40.0 / fs, 50.0 / fs)
    ctx.apply(k, pcm, frames, ch, bits, True)
    y0 = ctx.parked(frames, ch)
    p0 = ctx.peak()
    try:
        ctx.set_x_budget(1 << 20)          # 1 MiB -> ~16 chunks
        ctx.apply(k, pcm, frames, ch, bits, True)
        assert ctx.last_timing()["fir_launches"] > 4
        assert np.array_equal(ctx.parked(frames, ch), y0) and ctx.peak() == p0
    finally:
        ctx.set_x_budget(2 << 30)
        k.free()


# ------------------------------------------------------- synthetic generator ------

@pytest.mark.parametrize("bits,be,ch", [(16, True, 2), (24, False, 2), (24, False, 8), (32, False, 16)])
def test_device_synth_equals_oracle_synth(ctx, oracle_mod, bits, be, ch):
    import torch

    frames, first, rate = 20000, 123_456_789, 96000
    d = torch.empty(frames * ch * bits // 8, dtype=torch.uint8, device="cuda:0")
    ctx.synth_pcm_dev(0xF1F1F1, first, frames, ch, bits, be, rate, 1.0, d)
    ctx.synchronize()
    want = oracle_mod.synth_pcm(0xF1F1F1, first, frames, ch, bits, be, rate)
    assert np.array_equal(d.cpu().numpy(), want)


# ------------------------------------------------- full BASELINE sizes: properties --

def test_config1_full_size_windows_and_properties(ctx, oracle_mod):
    """Config 1 at its full size (60 s stereo 48 kHz 24-bit, 9601 taps), device-resident
    synthetic PCM: sampled windows (both file edges + random interior) against the oracle,
    plus DC rejection over the whole file."""
    import torch

    c = CONFIGS[1]
    frames, ch, bits, fs = c["frames"], c["channels"], c["bits"], c["fs"]
    fb = ch * bits // 8
    d_in = torch.empty(frames * fb, dtype=torch.uint8, device="cuda:0")
    ctx.synth_pcm_dev(0xF1F1F1, 0, frames, ch, bits, False, fs, 1.0, d_in)
    k = ctx.build_kernel(c["freq"] / fs, c["slope"] / fs)
    ctx.apply_dev(k, d_in, frames, ch, bits, False)
    pk = ctx.peak()
    y = ctx.parked(frames, ch)
    assert pk == float(np.abs(y).max()) and pk < 1.0
    taps = k.taps()
    H = k.half_len
    W = 4096
    rng = np.random.default_rng(1)
    starts = [0, frames - W] + [int(s) for s in rng.integers(H, frames - H - W, 6)]
    for s in starts:
        lo, hi = max(0, s - H), min(frames, s + W + H)
        seg = oracle_mod.synth_pcm(0xF1F1F1, lo, hi - lo, ch, bits, False, fs)
        x = oracle_mod.decode(seg, hi - lo, ch, bits, False)
        for cc in range(ch):
            xx = x[cc]
            # [lo, hi) is clipped only at the true file ends, where zero padding is right
            want = oracle_mod.fir_hi(xx, taps, s - lo, s - lo + W)[s - lo:s - lo + W]
            scale = oracle_mod.fir_abs_scale(xx, taps, s - lo, s - lo + W)[s - lo:s - lo + W]
            got = y[cc, s:s + W]
            assert np.all(np.abs(got - want) <= TOL * scale)
    # the 0.05 FS offset is gone in the steady state
    assert abs(y[:, H:frames - H].mean()) < 1e-6
    out = torch.empty_like(d_in)
    ctx.encode_dev(1.0, out)
    ctx.synchronize()
    want_pcm = oracle_mod.encode(y[:, :65536], 1.0, bits, False)
    n, mx = lsb_flips(out[:65536 * fb].cpu().numpy(), want_pcm, bits, False)
    assert mx <= 1 and n <= 2
    k.free()


# ------------------------------------------------------------- error paths --------

def test_error_behaviour(ctx):
    from audio_fir_filter_b200 import capi

    k = ctx.kernel_from_taps(np.array([0.25, 0.5, 0.25]))
    pcm = np.zeros(64, dtype=np.uint8)
    for kw in (dict(bits=12), dict(channels=0), dict(frames=-1)):
        a = dict(frames=8, channels=2, bits=16)
        a.update(kw)
        with pytest.raises(capi.FirGpuError) as e:
            ctx.apply(k, pcm, a["frames"], a["channels"], a["bits"], False)
        assert e.value.code == capi.ERR_INVALID
    with pytest.raises(capi.FirGpuError) as e:
        ctx.kernel_from_taps(np.array([0.5, 0.5]))      # even tap count: no centre tap
    assert e.value.code == capi.ERR_INVALID
    c2 = capi.Context(0)
    with pytest.raises(capi.FirGpuError) as e:
        c2.peak()                                        # nothing parked yet
    assert e.value.code == capi.ERR_STATE
    with pytest.raises(capi.FirGpuError) as e:
        c2.encode(1.0, pcm)
    assert e.value.code == capi.ERR_STATE
    c2.close()
    # empty payload: legal, produces nothing
    ctx.apply(k, pcm, 0, 2, 16, False)
    assert ctx.peak() == 0.0
    ctx.encode(1.0, pcm)
    k.free()


def test_halo_semantics_missing_halo_is_zero_extra_halo_is_ignored(ctx, oracle_mod):
    """fir_gpu_pcm.halo_left/right: real frames either side of the block; whatever is missing
    up to half_len is implicit zero (a true file edge, FilterCore.h:57-61,72-76), anything
    beyond half_len is accepted and ignored."""
    fs, ch, bits, be = 8000, 2, 24, False
    k = ctx.build_kernel(40.0 / fs, 50.0 / fs)                 # 641 taps, H = 320
    H = k.half_len
    total = 6000
    pcm = oracle_mod.synth_pcm(4, 0, total, ch, bits, be, fs)
    fb = ch * bits // 8
    taps = k.taps()
    x = oracle_mod.decode(pcm, total, ch, bits, be)
    s, e = 2000, 3504
    for hl, hr in [(H, H), (H + 123, H + 77), (100, 50), (0, 0), (H, 0)]:
        lo, hi = s - hl, e + hr
        ctx.apply(k, pcm[lo * fb:hi * fb], e - s, ch, bits, be, hl, hr)
        y = ctx.parked(e - s, ch)
        for c in range(ch):
            xx = np.zeros(total)
            a, b = max(lo, s - H), min(hi, e + H)              # what the device may use
            xx[a:b] = x[c, a:b]
            want = oracle_mod.fir_hi(xx, taps, s, e)[s:e]
            scale = oracle_mod.fir_abs_scale(xx, taps, s, e)[s:e]
            assert np.all(np.abs(y[c] - want) <= TOL * scale + 1e-300), (hl, hr)
    k.free()


def test_streamed_apply_progress_and_ranged_encode_are_bit_exact(ctx, oracle_mod):
    """fir_gpu_apply_begin/_feed/_end with the payload split at arbitrary byte positions ==
    fir_gpu_apply; the progress hook reports every chunk; fir_gpu_encode_range pieces ==
    fir_gpu_encode."""
    from audio_fir_filter_b200 import capi

    fs, ch, bits, be = 8000, 3, 24, True
    frames, hl, hr = 200_000, 100, 320
    k = ctx.build_kernel(40.0 / fs, 50.0 / fs)
    pcm = oracle_mod.synth_pcm(8, 0, hl + frames + hr, ch, bits, be, fs)
    ctx.apply(k, pcm, frames, ch, bits, be, hl, hr)
    y0, p0 = ctx.parked(frames, ch), ctx.peak()
    whole = np.empty(frames * ch * bits // 8, dtype=np.uint8)
    ctx.encode(0.9, whole)

    cuts = [0, 1, 7, 4096, 4099, 300_001, 300_001, 1_000_000, pcm.size]      # odd sizes, an empty piece
    pieces = [pcm[a:b].copy() for a, b in zip(cuts[:-1], cuts[1:])]
    seen = []
    ctx.set_progress(lambda d, t: seen.append((d, t)))
    try:
        ctx.apply_streamed(k, pieces, frames, ch, bits, be, hl, hr)
        assert ctx.peak() == p0
    finally:
        ctx.set_progress(None)
    assert np.array_equal(ctx.parked(frames, ch), y0)
    assert len(seen) >= 2 and seen[-1] == (frames, frames)
    assert all(a[0] < b[0] for a, b in zip(seen[:-1], seen[1:]))

    fb = ch * bits // 8
    out = np.zeros_like(whole)
    for f0 in range(0, frames, 65_536):
        n = min(65_536, frames - f0)
        ctx.encode_range(0.9, f0, n, out[f0 * fb:(f0 + n) * fb])
    assert np.array_equal(out, whole)

    # protocol errors
    fmt = ctx._fmt(1000, ch, bits, be)
    import ctypes as C
    L = capi.lib()
    assert L.fir_gpu_apply_feed(ctx._h, pcm.ctypes.data, 10) == capi.ERR_STATE          # nothing open
    assert L.fir_gpu_apply_begin(ctx._h, k._h, C.byref(fmt)) == capi.OK
    assert L.fir_gpu_apply_feed(ctx._h, pcm.ctypes.data, 1000 * fb + 1) == capi.ERR_INVALID   # too many bytes
    assert L.fir_gpu_apply_feed(ctx._h, pcm.ctypes.data, 500 * fb) == capi.OK
    assert L.fir_gpu_apply_end(ctx._h) == capi.ERR_STATE                                 # bytes missing
    with pytest.raises(capi.FirGpuError):
        ctx.encode_range(1.0, 1, 10, out)                                               # odd first frame
    k.free()


@pytest.mark.parametrize("bits,be,ch,frames,shift", [
    (24, False, 7, 1, 0), (24, True, 7, 2, 1), (16, False, 1, 3, 1), (32, True, 5, 17, 3), (24, False, 1, 33, 2),
    (16, True, 6, 4097, 1), (24, False, 2, 1365, 3), (32, False, 256, 40, 0), (24, True, 3, 70_001, 5),
])
def test_ragged_and_misaligned_payloads(ctx, oracle_mod, bits, be, ch, frames, shift):
    """Tiny files (shorter than the kernel: the reference's UB region, here the zero-padded
    formula), odd frame sizes, and payloads that start at any byte address -- the codec
    kernels' ragged head/tail paths must not touch a byte outside the payload."""
    import torch

    fs = 8000
    k = ctx.build_kernel(40.0 / fs, 50.0 / fs)                 # 641 taps
    n = frames * ch * bits // 8
    pcm = oracle_mod.synth_pcm(frames + ch, 0, frames, ch, bits, be, fs)
    want = oracle_mod.process(pcm, frames, ch, bits, be, 40.0 / fs, 50.0 / fs, True)
    # host path, misaligned host pointer
    hbuf = np.zeros(n + 64, dtype=np.uint8)
    hbuf[shift:shift + n] = pcm
    ctx.apply(k, hbuf[shift:shift + n], frames, ch, bits, be)
    y = ctx.parked(frames, ch)
    assert np.max(np.abs(y - want["y"])) <= 1e-12 * max(np.abs(want["y"]).max(), 1e-30)
    pk = ctx.peak()
    out = np.full(n + 64, 0xA5, dtype=np.uint8)
    ctx.encode(1.0 / pk, out[shift:shift + n])
    assert np.all(out[:shift] == 0xA5) and np.all(out[shift + n:] == 0xA5)      # guard bytes untouched
    nflip, mx = lsb_flips(out[shift:shift + n], want["pcm"], bits, be)
    assert mx <= 1 and nflip <= 2
    # device path, misaligned device pointers for both input and output
    d_in = torch.zeros(n + 64, dtype=torch.uint8, device="cuda:0")
    d_in[shift:shift + n] = torch.from_numpy(pcm).cuda()
    d_out = torch.full((n + 64,), 0xA5, dtype=torch.uint8, device="cuda:0")
    ctx.apply_dev(k, d_in.data_ptr() + shift, frames, ch, bits, be)
    assert ctx.peak() == pk
    ctx.encode_dev(1.0 / pk, d_out.data_ptr() + shift)
    ctx.synchronize()
    o = d_out.cpu().numpy()
    assert np.array_equal(o[shift:shift + n], out[shift:shift + n])
    assert np.all(o[:shift] == 0xA5) and np.all(o[shift + n:] == 0xA5)
    k.free()


def test_reference_make_test_settings_whole_path(ctx, oracle_mod):
    """`lowcut -f 440 -s 80 -n` -- the settings of the reference's own `make test`
    (Makefile:47) -- on a 44.1 kHz stereo 16-bit file: odd 4/bw (2205 -> M = 2206)."""
    from audio_fir_filter_b200 import FilterOptions, PcmInfo, process_pcm

    fs, ch, bits, frames = 44100, 2, 16, 150_000
    pcm = oracle_mod.synth_pcm(440, 0, frames, ch, bits, False, fs, gain=1.1)
    out = np.empty_like(pcm)
    r = process_pcm(ctx, pcm, PcmInfo(frames, ch, bits, False, float(fs)), FilterOptions(440.0, 80.0, True), out)
    assert r["taps"] == 2207
    want = oracle_mod.process(pcm, frames, ch, bits, False, 440.0 / fs, 80.0 / fs, True)
    assert abs(r["peak"] - want["peak"]) <= 1e-12 * want["peak"]
    nflip, mx = lsb_flips(out, want["pcm"], bits, False)
    assert mx <= 1 and nflip <= 2


def test_repeated_passes_do_not_leak_device_memory(oracle_mod):
    """300 apply/peak/encode passes of varying size on one context, with a progress hook and
    two kernels alive: results stay identical and the device's free memory does not drift."""
    import torch

    from audio_fir_filter_b200 import Context

    fs, ch, bits = 8000, 2, 16
    sizes = [4000, 70_000, 12_345, 1, 33_000]
    pcms = {n: oracle_mod.synth_pcm(n, 0, n, ch, bits, False, fs) for n in sizes}
    with Context(0) as cx:
        cx.set_progress(lambda d, t: None)
        k1, k2 = cx.build_kernel(40.0 / fs, 50.0 / fs), cx.build_kernel(100.0 / fs, 400.0 / fs)
        first = {}
        free0 = None
        for it in range(300):
            n = sizes[it % len(sizes)]
            k = k1 if it % 2 else k2
            out = np.empty_like(pcms[n])
            cx.apply(k, pcms[n], n, ch, bits, False)
            pk = cx.peak()
            cx.encode(1.0 / pk if pk > 0 else 1.0, out)
            key = (n, it % 2)
            if key in first:
                assert np.array_equal(out, first[key][0]) and pk == first[key][1]
            else:
                first[key] = (out, pk)
            if it == 20:
                free0 = torch.cuda.mem_get_info(0)[0]
        cx.set_progress(None)
        assert abs(torch.cuda.mem_get_info(0)[0] - free0) < (64 << 20)
        k1.free()
        k2.free()


def test_torch_interop_stream_and_device_peak_view(oracle_mod):
    """What bench.py and dist.py rely on: the context runs on a torch stream (torch events
    bracket its work), and the device-resident peak scalar is viewable by torch without a copy
    (the NCCL all-reduce runs on that view)."""
    import torch

    from audio_fir_filter_b200 import Context
    from audio_fir_filter_b200.dist import _DevScalar, allreduce_max_peak

    fs, ch, bits, frames = 8000, 2, 16, 300_000
    pcm = oracle_mod.synth_pcm(1, 0, frames, ch, bits, False, fs)
    d_in = torch.from_numpy(pcm).cuda()
    with Context(0) as cx:
        s = torch.cuda.Stream()
        cx.set_stream(s.cuda_stream)
        k = cx.build_kernel(40.0 / fs, 20.0 / fs)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        cx.apply_dev(k, d_in, frames, ch, bits, False)
        e1.record(s)
        s.synchronize()
        assert e0.elapsed_time(e1) >= cx.last_timing()["fir_ms"] > 0.0
        t = torch.as_tensor(_DevScalar(cx.peak_dev()), device="cuda:0")
        assert t.dtype == torch.float64 and float(t.item()) == cx.peak()
        assert allreduce_max_peak(cx) == cx.peak()           # world size 1: no collective
        cx.set_stream(None)                                   # back to the context's own stream
        cx.apply_dev(k, d_in, frames, ch, bits, False)
        assert cx.peak() == float(t.item())
        k.free()


@pytest.mark.parametrize("normalize,gain", [(False, 1.0), (True, 1.0), (False, 2.2), (True, 2.2)])
def test_one_call_process_equals_the_three_calls(ctx, oracle_mod, normalize, gain):
    """fir_gpu_process == fir_gpu_apply + fir_gpu_peak + scale rule + fir_gpu_encode, bit for bit --
    whether its bet (no -n and peak <= 1 -> scale 1, download under the FIR) wins (gain 1.0) or
    loses (gain 2.2 clips the input, the filtered peak exceeds 1 and everything is re-encoded)."""
    from audio_fir_filter_b200 import scale_for_peak

    fs, ch, bits, be, frames = 48000, 2, 24, False, 4_000_000
    if gain > 2.0:   # a full-scale square wave overshoots after the high-pass: filtered peak > 1
        sq = np.where((np.arange(frames) // 50) % 2 == 0, (1 << 23) - 1, -(1 << 23)).astype(np.int32)
        b = np.stack([(sq >> sh) & 0xFF for sh in (0, 8, 16)], axis=1).astype(np.uint8)
        pcm = np.ascontiguousarray(np.repeat(b[:, None, :], ch, axis=1).reshape(-1))
    else:
        pcm = oracle_mod.synth_pcm(6, 0, frames, ch, bits, be, fs)
    k = ctx.build_kernel(20.0 / fs, 100.0 / fs)
    ctx.apply(k, pcm, frames, ch, bits, be)
    pk = ctx.peak()
    sc = scale_for_peak(pk, normalize)
    want = np.empty_like(pcm)
    ctx.encode(sc, want)
    got = np.full_like(pcm, 0x5A)
    pk2, sc2 = ctx.process(k, pcm, frames, ch, bits, be, normalize, got)
    assert (pk2, sc2) == (pk, sc)
    assert (pk > 1.0) == (gain > 2.0)
    assert np.array_equal(got, want)
    t = ctx.last_timing()
    if not normalize:
        assert t["fir_launches"] >= 2          # the speculative path filters chunk by chunk
    k.free()


def test_out_of_memory_is_an_error_code_and_the_context_survives(ctx, oracle_mod):
    """A payload whose parked FP64 signal cannot fit (8 TB) fails with FIR_GPU_ERR_NOMEM before any
    kernel runs; the context keeps working afterwards."""
    import torch

    from audio_fir_filter_b200 import capi

    fs, ch, bits = 8000, 2, 16
    k = ctx.build_kernel(40.0 / fs, 50.0 / fs)
    small = oracle_mod.synth_pcm(2, 0, 10_000, ch, bits, False, fs)
    d = torch.from_numpy(small).cuda()
    with pytest.raises(capi.FirGpuError) as e:
        ctx.apply_dev(k, d, 1 << 39, ch, bits, False)          # 2^39 frames x 2 ch x 8 B = 8 TB
    assert e.value.code == capi.ERR_NOMEM
    with pytest.raises(capi.FirGpuError) as e:
        ctx.peak()                                              # nothing got parked
    assert e.value.code == capi.ERR_STATE
    out = np.empty_like(small)
    pk, sc = ctx.process(k, small, 10_000, ch, bits, False, True, out)
    want = oracle_mod.process(small, 10_000, ch, bits, False, 40.0 / fs, 50.0 / fs, True)
    assert abs(pk - want["peak"]) <= 1e-12 * want["peak"] and np.array_equal(out, want["pcm"])
    k.free()


def test_failed_create_leaves_no_device_memory_behind():
    """fir_gpu_create that fails after its streams, events and device buffers exist (test hook of
    fir_gpu_dev.h) must tear the half-built context down: free device memory is unchanged and
    the next create works (VERDICT r1 #9 / ADVICE r1)."""
    import torch

    from audio_fir_filter_b200 import capi

    capi.Context(0).close()                                # CUDA context and lazy modules are up
    torch.cuda.synchronize()
    free0, _ = torch.cuda.mem_get_info(0)
    for _ in range(20):
        capi.lib().fir_gpu_test_fail_next_create(1)
        with pytest.raises(capi.FirGpuError) as e:
            capi.Context(0)
        assert e.value.code == capi.ERR_CUDA and "injected" in str(e.value)
    free1, _ = torch.cuda.mem_get_info(0)
    assert free0 - free1 < (2 << 20), (free0, free1)       # 20 leaked contexts would hold >= 20 x 2 MiB granules
    with capi.Context(0) as c:                             # and the library is still usable
        assert c.fp64_peak(1, 0.01) > 1.0


@pytest.mark.parametrize("bits", [16, 24, 32])
def test_encode_saturates_like_the_oracle_for_absurd_gains(ctx, oracle_mod, bits):
    """The encoder converts with saturation and clamps as an integer; the oracle clamps in floating
    point and then rounds.  They must agree for any gain, however far beyond full scale."""
    fs, frames, ch = 8000, 4096, 2
    pcm = oracle_mod.synth_pcm(77, 0, frames, ch, bits, False, fs)
    k = ctx.build_kernel(40.0 / fs, 50.0 / fs)
    ctx.apply(k, pcm, frames, ch, bits, False)
    y = ctx.parked(frames, ch)
    out = np.empty_like(pcm)
    for scale in (1.0, 7.5, 1e6, 1e30, 1e290):
        ctx.encode(scale, out)
        want = oracle_mod.encode(y, scale, bits, False)
        assert np.array_equal(out, want), scale
    k.free()


def test_peak_allreduce_entry_point_on_one_device(ctx, oracle_mod):
    """fir_gpu_allreduce_peak with a single block is that block's peak (no collective); two blocks
    on ONE device are refused -- a sample block per device is the contract."""
    from audio_fir_filter_b200 import Context, capi

    fs, frames, ch, bits = 8000, 20_000, 2, 16
    pcm = oracle_mod.synth_pcm(3, 0, frames, ch, bits, False, fs)
    k = ctx.build_kernel(40.0 / fs, 50.0 / fs)
    ctx.apply(k, pcm, frames, ch, bits, False)
    assert capi.allreduce_peak([ctx]) == ctx.peak()
    other = Context(0)
    other.apply(k, pcm, frames, ch, bits, False)
    with pytest.raises(capi.FirGpuError) as e:
        capi.allreduce_peak([ctx, other])
    assert e.value.code in (capi.ERR_INVALID, capi.ERR_STATE)      # INVALID: same device (STATE only if NCCL is absent)
    fresh = Context(0)
    with pytest.raises(capi.FirGpuError) as e:
        capi.allreduce_peak([fresh])
    assert e.value.code == capi.ERR_STATE                            # nothing parked
    fresh.close()
    other.close()
    k.free()
