#!/usr/bin/env python
"""GPU tool: HBM throughput of pcm_decode_kernel / pcm_encode_kernel for the PCM formats of the
BASELINE configs, over tile sizes and CTA widths (fir_gpu_set_codec_geometry).  The FIR in
between is an identity kernel (one tap), so a pass costs almost nothing but the two codec
launches; times are the library's own CUDA-event spans.  Usage:
    python tools/sweep_codec.py [--mb 400] [--reps 5] [--geoms 8192x256,16384x256,...]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from audio_fir_filter_b200 import capi  # noqa: E402
from bench import SEED, peak_hbm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mb", type=float, default=400.0, help="PCM megabytes per pass")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--geoms", default="0x256x50,8192x256x50,12288x256x50,16384x256x50,16384x128x50,32768x256x50,0x256x-1",
                help="tile bytes x threads x shared-memory carveout percent (-1 = driver's choice)")
ap.add_argument("--formats", default="", help="comma-separated substrings of the format names to run (default: all)")
a = ap.parse_args()
FORMATS = [("cfg2 16BE x2", 2, 16, True, 44100), ("cfg1/4 24LE x2", 2, 24, False, 48000),
           ("cfg3 24LE x8", 8, 24, False, 96000), ("cfg5 32LE x16", 16, 32, False, 192000),
           ("mono 16LE", 1, 16, False, 44100)]
ctx = capi.Context(0)
k = ctx.kernel_from_taps(np.array([1.0]))
peak = peak_hbm()
best = {}
for name, ch, bits, be, fs in FORMATS:
    if a.formats and not any(s in name for s in a.formats.split(",")):
        continue
    fb = ch * bits // 8
    frames = int(a.mb * 1e6 / fb) & ~1023
    d_in = torch.empty(frames * fb, dtype=torch.uint8, device="cuda:0")
    d_out = torch.empty_like(d_in)
    ctx.synth_pcm_dev(SEED, 0, frames, ch, bits, be, fs, 1.0, d_in)
    ctx.synchronize()
    dec_bytes, enc_bytes = frames * fb + frames * ch * 8, frames * ch * 8 + frames * fb
    ref = None
    for g in a.geoms.split(","):
        tile, nt, carve = (int(v) for v in (g.split("x") + ["50"])[:3])
        ctx.set_codec_geometry(tile, nt, carve)
        dec, enc = [], []
        for _ in range(a.reps + 1):
            ctx.apply_dev(k, d_in, frames, ch, bits, be)
            ctx.encode_dev(1.0, d_out)
            t = ctx.last_timing()
            dec.append(t["decode_ms"])
            enc.append(t["encode_ms"])
        if ref is None:
            ref = d_out.clone()
            assert torch.equal(ref, d_in), "identity kernel + scale 1 must reproduce the PCM"
        assert torch.equal(ref, d_out), f"geometry {g} changes the bytes"
        r = {"format": name, "geom": g, "decode_gbs": dec_bytes / min(dec[1:]) / 1e6, "encode_gbs": enc_bytes / min(enc[1:]) / 1e6,
             "decode_ms": min(dec[1:]), "encode_ms": min(enc[1:])}
        r["decode_frac"], r["encode_frac"] = r["decode_gbs"] / peak, r["encode_gbs"] / peak
        print(json.dumps(r), flush=True)
        b = best.setdefault(name, {})
        for key in ("decode_gbs", "encode_gbs"):
            if key not in b or r[key] > b[key][0]:
                b[key] = (r[key], g)
    del d_in, d_out
    torch.cuda.empty_cache()
print(json.dumps({"best": best, "peak_gbs": peak}))
