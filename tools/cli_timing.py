#!/usr/bin/env python
"""GPU-box tool: wall time of the C++ host (`host/lowcut`) on real files in /tmp --
one config-4-sized WAV (5 min stereo 48 kHz 24-bit), the config-2 AIFF, and a small batch.
The input PCM comes from the library's own device generator (fir_gpu_synth_pcm_dev)."""
import os
import subprocess
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch  # noqa: E402

from audio_fir_filter_b200 import capi  # noqa: E402
from audio_fixtures import aiff_bytes, wav_bytes  # noqa: E402


def synth(seed, frames, ch, bits, be, rate):
    with capi.Context(0) as ctx:
        d = torch.empty(frames * ch * bits // 8, dtype=torch.uint8, device="cuda:0")
        ctx.synth_pcm_dev(seed, 0, frames, ch, bits, be, rate, 1.0, d)
        ctx.synchronize()
        return d.cpu().numpy().tobytes()


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LOWCUT = os.path.join(ROOT, "host", "lowcut")
subprocess.run(["make", "-C", os.path.join(ROOT, "host")], check=True, capture_output=True)


def timed(*args):
    t0 = time.perf_counter()
    r = subprocess.run([LOWCUT, *map(str, args)], capture_output=True, text=True)
    dt = time.perf_counter() - t0
    assert r.returncode == 0, r.stderr
    return dt, r.stdout


with tempfile.TemporaryDirectory(dir=os.environ.get("TIMING_DIR", "/tmp")) as d:
    pcm = synth(1, 14_400_000, 2, 24, False, 48000)
    w = os.path.join(d, "cfg4.wav")
    open(w, "wb").write(wav_bytes(pcm, 2, 24, 48000))
    dt, out = timed("-v", "-f", 20, "-s", 20, w, os.path.join(d, "cfg4_out.wav"))
    print(f"cfg4 file ({len(pcm) / 1e6:.0f} MB, 9601 taps): {dt:.3f} s wall")
    print("\n".join(l for l in out.splitlines() if "device time" in l or "peak" in l))
    files = []
    NB = int(os.environ.get("BATCH", "32"))
    for i in range(NB):
        p = os.path.join(d, f"b{i}.wav")
        os.link(w, p)
        files.append(p)
    dt, _ = timed("-f", 20, "-s", 20, *files, os.path.join(d, "outdir"))
    print(f"batch of {NB} such files: {dt:.3f} s wall ({NB * 28.8 / dt:.0f} MSamples/s incl. file I/O and start-up)")
    pcm = synth(2, 26_460_000, 2, 16, True, 44100)
    a = os.path.join(d, "cfg2.aif")
    open(a, "wb").write(aiff_bytes(pcm, 2, 16, 44100.0))
    dt, out = timed("-v", "-n", "-f", 30, "-s", 10, a, os.path.join(d, "cfg2_out.aif"))
    print(f"cfg2 file ({len(pcm) / 1e6:.0f} MB, 17641 taps, -n): {dt:.3f} s wall ({52.92 / dt:.0f} MSamples/s incl. file I/O)")
    print("\n".join(l for l in out.splitlines() if "device time" in l or "peak" in l))
