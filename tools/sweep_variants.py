#!/usr/bin/env python
"""GPU tool: FP64 probes and a sweep of the FIR kernel variants on device-resident
synthetic PCM (CUDA-event time of the FIR launches only).  Usage:
    python tools/sweep_variants.py [--config 2] [--frames N] [--reps 3]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
# the product library carries two FIR shapes; the sweep needs the -DFIR_ALL_VARIANTS build
SWEEP_LIB = os.path.join(ROOT, "audio_fir_filter_b200", "libfir_gpu_sweep.so")
if "FIR_GPU_LIB" not in os.environ and os.path.exists(SWEEP_LIB):
    os.environ["FIR_GPU_LIB"] = SWEEP_LIB
import torch  # noqa: E402

from audio_fir_filter_b200 import capi  # noqa: E402
from bench import CONFIGS, SEED, algorithmic_flop  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=2)
ap.add_argument("--frames", type=int, default=0)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--variants", default="")
ap.add_argument("--slope", type=float, default=0.0, help="override the config's slope (tap count = 4*fs/slope + 1)")
a = ap.parse_args()
cfg = CONFIGS[a.config]
frames = a.frames or cfg["frames"]
fs, ch, bits, be = cfg["fs"], cfg["channels"], cfg["bits"], cfg["be"]
ctx = capi.Context(0)
print(json.dumps({"dfma_tflops": ctx.fp64_peak(0, 0.3), "dmma_tflops": ctx.fp64_peak(1, 0.3)}), flush=True)
slope = a.slope or cfg["slope"]
k = ctx.build_kernel(cfg["freq"] / fs, slope / fs)
taps = k.num_taps
d = torch.empty(frames * ch * bits // 8, dtype=torch.uint8, device="cuda:0")
ctx.synth_pcm_dev(SEED, 0, frames, ch, bits, be, fs, 1.0, d)
flop = algorithmic_flop(frames, ch, taps, 0, 0)
names = capi.variant_names()
sel = [int(v) for v in a.variants.split(",")] if a.variants else range(len(names))
for v in sel:
    ctx.set_variant(v)
    best = None
    for _ in range(a.reps + 1):
        ctx.apply_dev(k, d, frames, ch, bits, be)
        t = ctx.last_timing()
        best = t["fir_ms"] if best is None else min(best, t["fir_ms"])
    print(json.dumps({"variant": v, "name": names[v], "taps": taps, "fir_ms": best, "tflops": flop / best / 1e9,
                      "hbm_gbs": 16.0 * frames * ch / best / 1e6, "decode_ms": t["decode_ms"], "peak": ctx.peak()}),
          flush=True)
