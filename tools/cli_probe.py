#!/usr/bin/env python
"""GPU-box tool: `host/lowcut -v` on one config-4-sized WAVE (86 MB) twice, with the verbose
time stamps -- where the CLI's wall time goes (context bring-up vs. the file itself)."""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from audio_fir_filter_b200 import capi
from audio_fixtures import wav_bytes
with capi.Context(0) as ctx:
    n = 14_400_000
    d = torch.empty(n * 6, dtype=torch.uint8, device="cuda:0")
    ctx.synth_pcm_dev(1, 0, n, 2, 24, False, 48000, 1.0, d); ctx.synchronize()
    pcm = d.cpu().numpy().tobytes()
with tempfile.TemporaryDirectory(dir="/tmp") as t:
    w = os.path.join(t, "a.wav"); open(w, "wb").write(wav_bytes(pcm, 2, 24, 48000))
    for i in range(2):
        t0 = time.perf_counter()
        r = subprocess.run([os.path.join(ROOT, "host", "lowcut"), "-v", "-O", "-f", "20", "-s", "20", w, os.path.join(t, "o.wav")], capture_output=True, text=True)
        print("wall", round(time.perf_counter() - t0, 3)); print(r.stdout); print(r.stderr[-300:])
