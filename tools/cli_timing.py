#!/usr/bin/env python
"""GPU-box tool: what a `lowcut` user waits for.  Wall time of the C++ host (`host/lowcut`) on
real files in a tmpfs directory (TIMING_DIR, default /dev/shm): config 1, config 2 (-n), one
config-4 file, and the config-4 batch (BATCH files, default 32; 256 = the whole of config 4)
on -g GPUS devices -- each with the start-up / read+filter / write break-down the -v time
stamps give.  The input PCM comes from the library's own device generator."""
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from audio_fir_filter_b200 import capi  # noqa: E402
from audio_fixtures import aiff_bytes, wav_bytes  # noqa: E402

LOWCUT = os.path.join(ROOT, "host", "lowcut")
subprocess.run(["make", "-C", os.path.join(ROOT, "host")], check=True, capture_output=True)
GPUS = int(os.environ.get("GPUS", "0"))          # batch case only: -g GPUS (0 = let lowcut choose, as for the single files)
ONLY = os.environ.get("ONLY", "")                # "batch": skip the single-file cases
NB = int(os.environ.get("BATCH", "32"))
REPS = int(os.environ.get("REPS", "3"))


def synth(seed, frames, ch, bits, be, rate):
    with capi.Context(0) as ctx:
        d = torch.empty(frames * ch * bits // 8, dtype=torch.uint8, device="cuda:0")
        ctx.synth_pcm_dev(seed, 0, frames, ch, bits, be, rate, 1.0, d)
        ctx.synchronize()
        return d.cpu().numpy().tobytes()


def timed(*args, gpus=0):
    extra = ["-g", str(gpus)] if gpus else []
    best = None
    for _ in range(REPS):
        t0 = time.perf_counter()
        r = subprocess.run([LOWCUT, "-O", *extra, *map(str, args)], capture_output=True, text=True)
        dt = time.perf_counter() - t0
        assert r.returncode == 0, r.stderr
        if best is None or dt < best[0]:
            best = (dt, r.stdout)
    return best


def stamps(out):
    """-v time stamps: seconds since program start for the milestones of a single file."""
    s = {}
    for key, pat in (("devices_counted", r"\[\s*([\d.]+) s since start\] devices counted"),
                     ("contexts_ready", r"\[\s*([\d.]+) s since start\] \d+ GPU context"),
                     ("done", r"\[\s*([\d.]+) s since start\] done")):
        m = re.search(pat, out)
        if m:
            s[key] = float(m.group(1))
    for key, pat in (("kernel_built", r"\[\s*([\d.]+) s\] kernel built"), ("filtered", r"\[\s*([\d.]+) s\] filtered"),
                     ("output_created", r"\[\s*([\d.]+) s\] output file created"),
                     ("written", r"\[\s*([\d.]+) s\] encoded and written")):
        m = re.search(pat, out)
        if m:
            s["file_" + key] = float(m.group(1))
    m = re.search(r"device time: (.*)", out)
    if m:
        s["device"] = m.group(1)
    return s


def report(name, wall, out, msamples):
    s = stamps(out)
    line = {"case": name, "wall_s": round(wall, 3), "msamples_per_s_wall": round(msamples / wall, 1), **s}
    if "contexts_ready" in s and "done" in s:
        line["breakdown"] = {
            "start_up_s (exec, CUDA init, device count, context)": s["contexts_ready"],
            "read+upload+filter_s": s.get("file_filtered"),
            "create_output+encode+download+write_s": round(s.get("file_written", 0) - s.get("file_filtered", 0), 3),
            "exit_s": round(wall - s["done"], 3),
        }
    print(json.dumps(line), flush=True)


with tempfile.TemporaryDirectory(dir=os.environ.get("TIMING_DIR", "/dev/shm")) as d:
    if ONLY != "batch":
        pcm = synth(3, 2_880_000, 2, 24, False, 48000)
        w1 = os.path.join(d, "cfg1.wav")
        open(w1, "wb").write(wav_bytes(pcm, 2, 24, 48000))
        dt, out = timed("-v", "-f", 20, "-s", 20, w1, os.path.join(d, "cfg1_out.wav"))
        report("cfg1: 60 s stereo 48 kHz 24-bit WAV, 9601 taps", dt, out, 5.76)

        pcm = synth(2, 26_460_000, 2, 16, True, 44100)
        a = os.path.join(d, "cfg2.aif")
        open(a, "wb").write(aiff_bytes(pcm, 2, 16, 44100.0))
        dt, out = timed("-v", "-n", "-f", 30, "-s", 10, a, os.path.join(d, "cfg2_out.aif"))
        report("cfg2: 10 min stereo 44.1 kHz 16-bit BE AIFF, -n, 17641 taps", dt, out, 52.92)

    pcm = synth(1, 14_400_000, 2, 24, False, 48000)
    w = os.path.join(d, "cfg4.wav")
    open(w, "wb").write(wav_bytes(pcm, 2, 24, 48000))
    if ONLY != "batch":
        dt, out = timed("-v", "-f", 20, "-s", 20, w, os.path.join(d, "cfg4_out.wav"))
        report("cfg4: one 5 min stereo 48 kHz 24-bit WAV, 9601 taps", dt, out, 28.8)

    files = []
    for i in range(NB):
        p = os.path.join(d, f"b{i}.wav")
        os.link(w, p)
        files.append(p)
    # the ceiling of the file system itself: the same number of files and bytes, created by as many threads as
    # lowcut has lanes (4 per GPU), doing nothing else
    probe = os.path.join(d, "tmpfs_write_probe")
    subprocess.run(["g++", "-O2", "-pthread", "-o", probe, os.path.join(ROOT, "tools", "tmpfs_write_probe.cpp")], check=True)
    os.mkdir(os.path.join(d, "probe"))
    ngpu = GPUS or capi.device_count()
    for pre in (0, 1):
        r = subprocess.run([probe, os.path.join(d, "probe"), str(4 * ngpu), str(NB), "86", str(pre)], capture_output=True, text=True)
        print(r.stdout.strip(), flush=True)
    os.environ["LOWCUT_TRACE"] = "1"        # per-process phase totals (thread-seconds) after the batch
    # lowcut's own choice of GPUs first (sqrt(FIR seconds / 0.5 s per GPU)), then every count asked for in GPUS_SWEEP
    sweep = [GPUS] + [int(v) for v in os.environ.get("GPUS_SWEEP", "").split(",") if v]
    for want in sweep:
        dt, out = timed("-f", 20, "-s", 20, *files, os.path.join(d, "outdir"), gpus=want)
        for l in out.splitlines():
            if "trace (pid" in l:
                print(l.strip(), flush=True)
        m = re.search(r"Using up to (\d+) GPU", out)
        g = int(m.group(1)) if m else 1
        print(json.dumps({"case": f"cfg4 batch: {NB} such files ({NB * 86.4 / 1e3:.1f} GB in, as much out) to a directory",
                          "gpus": g, "chosen_by": "lowcut" if not want else f"-g {want}",
                          "processes": "one per GPU" if "one worker process per GPU" in out else "one",
                          "wall_s": round(dt, 3), "msamples_per_s_wall": round(NB * 28.8 / dt, 1),
                          "fir_device_s_per_gpu": round(NB * 0.0153 / g, 3)}), flush=True)
