#!/usr/bin/env python
"""bench.py -- the lowcut hot path on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C] [--impl ours|reference]

A *step* is one pass of the hot path over one synthetic payload of the shape
BASELINE.json names (default: configs[1], the 10 min stereo 44.1 kHz 16-bit
big-endian AIFF with ``-f 30 -s 10 -n``):

    build_kernel (once, outside the loop: it depends only on the file's rate)
    apply   = PCM decode -> FP64 FIR over every channel (+ fused peak)
    peak    = 8-byte readback (N>1: NCCL all-reduce MAX of that scalar first)
    encode  = scale, round, clip, endian, interleave

``value``   : filtered output MSamples/s of the whole job with the PCM resident in HBM
              (device timing, CUDA events on the launching stream, max over ranks).
``e2e``     : the same metric through the host-buffer C-ABI (fir_gpu_process, or
              fir_gpu_apply / fir_gpu_peak / fir_gpu_encode at N>1): pinned host PCM in,
              pinned host PCM out, both copies inside the timed region.  ``e2e.pipelined``
              is the same work with two contexts per rank, file k's way out running under
              file k+1's FIR; ``e2e.copy_ceiling`` is what plain pinned copies reach here.
``roofline``: the FIR kernel's algorithmic FP64 FLOP over its CUDA-event time, against
              the FP64 tensor-pipe peak measured live by a register-resident probe.
``parity``  : every rank checks windows of what it just produced (its first and last
              frames = both sides of every block seam and the true file edges, plus
              random interior positions) against the CPU oracle: the parked FP64 signal
              at the D3 tolerance, the encoded PCM (device arm and host arm) for 1-LSB
              flips, the taps, the peak.  A failure makes the process exit non-zero.
``cli``     : N=1: the shipped C++ host (host/lowcut) on a tmpfs file of the workload's shape -- wall
              clock with its start-up / filter / write / exit break-down, output byte-identical to what
              this process's own context makes of the same payload.  N>1: ``cli_block_mode`` instead
              (lowcut -g N against -g 1, byte-identical; seam windows against the oracle).
``also``    : one record per other BASELINE config (N=1: configs 1, 4, a config-5 slice and
              config 3 IN FULL; N=8: config 3 split into sample blocks), each with its own
              parity block.
N > 1       : ONE long file of N x the payload, sample-block sharded with a (taps-1)
              halo; the only collective is the all-reduce MAX of the peak.
              ``--mode batch`` instead gives every rank its own files (no collective).

--impl reference times the CPU restatement of the reference's multithreaded path
(oracle/_ref = the reference's FilterCore.h compiled in place, else the oracle
port) on the host cores.  oracle/ is used here as the CPU baseline and as the parity
CHECKER only; nothing measured on the GPU arm touches it.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import shutil
import statistics
import struct
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SEED = 0x00F1F1F1
TOL = 1e-12          # decision D3: |y_gpu - y_oracle| <= TOL * sum_k |h_k * x_k|
METRIC = "filtered output MSamples/s (FP64 direct FIR low-cut, PCM in -> PCM out)"

CONFIGS = {
    1: dict(name="cfg1: 60 s stereo 48 kHz 24-bit LE WAV, -f 20 -s 20", fs=48000, freq=20.0, slope=20.0, channels=2,
            bits=24, be=False, normalize=False, frames=2_880_000),
    2: dict(name="cfg2: 10 min stereo 44.1 kHz 16-bit BE AIFF, -f 30 -s 10 -n", fs=44100, freq=30.0, slope=10.0,
            channels=2, bits=16, be=True, normalize=True, frames=26_460_000),
    3: dict(name="cfg3: 1 h 8-ch 96 kHz 24-bit LE WAV, -f 10 -s 2", fs=96000, freq=10.0, slope=2.0, channels=8,
            bits=24, be=False, normalize=False, frames=345_600_000),
    4: dict(name="cfg4: one 5 min stereo 48 kHz 24-bit LE WAV of the 256-file batch, -f 20 -s 20", fs=48000, freq=20.0,
            slope=20.0, channels=2, bits=24, be=False, normalize=False, frames=14_400_000),
    5: dict(name="cfg5 slice: 16-ch 192 kHz 32-bit LE WAV, -f 15 -s 5 -n, 2 min per GPU", fs=192000, freq=15.0,
            slope=5.0, channels=16, bits=32, be=False, normalize=True, frames=23_040_000),
    # config 5 IN FULL: only with --gpus 8 (one eighth of the 8 h file per GPU: 44 GB of PCM in place + 88 GB parked)
    6: dict(name="cfg5 FULL: 8 h 16-ch 192 kHz 32-bit LE WAV (353.9 GB), -f 15 -s 5 -n, one eighth per GPU", fs=192000,
            freq=15.0, slope=5.0, channels=16, bits=32, be=False, normalize=True, frames=691_200_000),
}


def kernel_order(bw_norm: float) -> int:
    m = int(round(4.0 / bw_norm))
    return m + (m & 1)


def algorithmic_flop(frames: int, channels: int, taps: int, halo_l: int, halo_r: int) -> float:
    """2 FLOP per in-range tap: taps*frames minus the taps that fall off
    a true file edge (FilterCore.h:58,73)."""
    H = (taps - 1) // 2
    missing = 0
    for halo in (halo_l, halo_r):
        m = max(0, H - halo)              # output n (from that edge) misses max(0, m - n) taps
        k = min(m, frames)
        missing += k * m - k * (k - 1) // 2
    return 2.0 * channels * (taps * frames - missing)


def config_dict(cfg: dict, world: int, mode: str, strong: bool = False) -> dict:
    """The `config` object of the JSON line -- built by BOTH arms from the same arguments, so the
    reference arm and the GPU arm name the same workload key for key."""
    fs = cfg["fs"]
    taps = kernel_order(cfg["slope"] / fs) + 1
    name = cfg["name"]
    if world > 1 and strong:
        name += f"; N={world}: the one file split into {world} sample blocks with {taps - 1}-frame halo + all-reduce MAX of the peak"
    elif world > 1 and mode == "block":
        name += (f"; N={world}: one file {world}x as long, sample-block sharded with {taps - 1}-frame halo + NCCL "
                 f"all-reduce MAX of the peak")
    elif world > 1:
        name += f"; N={world}: one such file per rank (batch mode)"
    frames_per_gpu = cfg["frames"] // world if strong else cfg["frames"]
    pcm = frames_per_gpu * cfg["channels"] * cfg["bits"] // 8
    return {
        "workload": name, "taps": taps, "frames_per_gpu": frames_per_gpu, "channels": cfg["channels"],
        "bits": cfg["bits"], "big_endian": cfg["be"], "normalize": cfg["normalize"], "sample_rate": fs,
        "l2": "inputs larger than L2 (PCM + FP64 planes per step >> 126 MB)" if pcm + 16 * frames_per_gpu *
              cfg["channels"] > 200e6 else "working set may fit L2; see DESIGN.md",
    }


def bind_to_gpu_numa_node(index: int):
    """Pin this rank to the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned host
    buffer is allocated (first touch puts the pages there): with 8 ranks moving 100 MB each way
    per step, buffers on the far socket halve the PCIe rates.  Best effort; returns a note."""
    try:
        out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(index)],
                             capture_output=True, text=True, timeout=20).stdout.strip()
        bus = out.lower()
        if bus.startswith("00000000:"):
            bus = "0000:" + bus.split(":", 1)[1]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return "numa: single node"
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"numa: node {node}, {len(cpus)} cpus"
        return "numa: no allowed cpu on the GPU's node"
    except Exception as e:  # noqa: BLE001
        return f"numa: not bound ({type(e).__name__})"


# ------------------------------------------------------------------ clocks ----------

class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        import threading

        self.proc = None
        self.lines = []
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line))

    def count_since(self, t0: float) -> int:
        return sum(1 for t, _ in self.lines if t >= t0)

    def stop(self, t0: float = 0.0, t1: float = float("inf")) -> dict:
        """Statistics over the samples taken between t0 and t1 (the timed region) when there
        are at least three, else over every sample drawing at least half the peak power."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.thread.join(timeout=2)
        inside = [l for t, l in self.lines if t0 <= t <= t1 + 0.12]
        window = "timed region" if len(inside) >= 3 else "all samples under load"
        out = "".join(inside if len(inside) >= 3 else [l for _, l in self.lines])
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.splitlines():
            p = [s.strip() for s in line.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
                pw.append(float(p[2]))
            except ValueError:
                continue
            for n, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        top = max(sm)
        load = [s for s, w in zip(sm, pw) if w >= 0.5 * max(pw)] or sm
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "sm_mhz_min": min(load), "sm_mhz_peak": top,
                "power_w_max": max(pw), "samples": len(sm), "window": window, "reasons": sorted(reasons)}


# ------------------------------------------------------------- CPU baseline ----------

def reference_threads(cores: int) -> int:
    """main.cp:75-76: floor(0.7 * hardware_concurrency), 4 when that is 0."""
    return int(math.floor(cores * 0.7)) or 4


def cpu_reference_pass(cfg: dict, frames: int, threads: int) -> dict:
    """ONE pass of the reference's CPU path over the first `frames` frames of the workload,
    PCM in -> PCM out: decode to float32 (readAll, ProcessFile.cp:41), the per-channel thread
    fan-out around FilterCore.h apply_filter_range (ProcessFile.cp:57-87; channels sequential),
    the peak loop (:92-96), the scale rule (:98) and normalise + encode (:100,117).  The FIR is
    oracle/_ref (the reference's FilterCore.h compiled in place) where it was built, else the
    oracle port; decode / encode are the oracle's float32 restatements of c_lib.  Returns the
    wall seconds of the whole pass and of the FIR alone."""
    import oracle

    fs, ch = cfg["fs"], cfg["channels"]
    taps = oracle.build_lowcut(cfg["freq"] / fs, cfg["slope"] / fs)
    use_ref = oracle.ref_lib() is not None
    pcm = oracle.synth_pcm(SEED, 0, frames, ch, cfg["bits"], cfg["be"], fs)   # input creation: untimed
    t0 = time.perf_counter()
    x = oracle.decode(pcm, frames, ch, cfg["bits"], cfg["be"], dtype=np.float32)
    t1 = time.perf_counter()
    y = np.empty_like(x)
    for c in range(ch):                                   # channels sequential (ProcessFile.cp:57)
        y[c] = (oracle.ref_filter_channel_threads(x[c], taps, threads) if use_ref
                else oracle.fir_f32(x[c], taps, threads=threads))
    t2 = time.perf_counter()
    peak = max(float(np.abs(y[c]).max()) for c in range(ch))                  # ProcessFile.cp:92-96
    scale = 1.0 / peak if (peak > 1.0 or cfg["normalize"]) and peak > 0 else 1.0
    oracle.encode_f32(y, scale, cfg["bits"], cfg["be"])
    t3 = time.perf_counter()
    return {"total_s": t3 - t0, "fir_s": t2 - t1, "decode_s": t1 - t0, "peak_encode_s": t3 - t2, "frames": frames,
            "kind": "reference" if use_ref else "port", "taps": int(taps.size),
            "simd": "AVX-512" if use_ref and "avx512" in getattr(oracle.ref_lib(), "_path", "") else "AVX2/FMA"}


def cpu_reference_rate(cfg: dict, budget_s: float, threads: int | None = None, frames: int | None = None):
    """A bounded sample of the workload through cpu_reference_pass -> (MSamples/s, description, seconds)."""
    import oracle

    oracle.build()
    threads = threads or len(os.sched_getaffinity(0))
    ch = cfg["channels"]
    taps = kernel_order(cfg["slope"] / cfg["fs"]) + 1
    if frames is None:
        n0 = max(4 * taps, 1 << 15)                       # calibrate on a short file, then size the sample
        r = cpu_reference_pass(cfg, n0, threads)
        rate = n0 * ch / r["total_s"]
        frames = int(min(cfg["frames"], max(n0, rate * budget_s / ch)))
    r = cpu_reference_pass(cfg, frames, threads)
    t = r["total_s"]
    msps = frames * ch / t / 1e6
    whole = frames == cfg["frames"]
    desc = {
        "kind": r["kind"],
        "cores": threads,
        "sample": (("the WHOLE file: " if whole else "first ") + f"{frames} frames x {ch} ch of the workload "
                   f"({frames * ch / 1e6:.2f} MSamples, {t:.1f} s), same {taps}-tap kernel, PCM in -> PCM out: decode "
                   f"{r['decode_s'] * 1e3:.0f} ms + FIR {r['fir_s']:.2f} s + peak/normalise/encode "
                   f"{r['peak_encode_s'] * 1e3:.0f} ms; FIR = " +
                   ("reference FilterCore.h apply_filter_range compiled in place (c_lib's fms()/taps behind interface "
                    "shims)" if r["kind"] == "reference" else "oracle port of FilterCore.h (ref_f32 mode)") +
                   f", float32 buffers, {threads} std::threads per channel as ProcessFile.cp:64-83, -O3 {r['simd']}"),
        "fir_share": r["fir_s"] / t,
    }
    return msps, desc, t


def cpu_rows(cfg: dict, budget_s: float) -> list:
    """both thread counts -- every core, and the reference's default
    floor(0.7 * cores) (main.cp:75)."""
    cores = len(os.sched_getaffinity(0))
    rows = []
    for th, b in ((cores, budget_s), (reference_threads(cores), max(3.0, budget_s / 3))):
        v, d, t = cpu_reference_rate(cfg, b, th)
        rows.append({"threads": th, "value": v, "unit": "MSamples/s", "seconds": t,
                     "row": "all cores" if th == cores else "reference default floor(0.7*cores), main.cp:75",
                     "sample": d["sample"], "fir_share": d["fir_share"], "kind": d["kind"]})
    return rows


# ------------------------------------------------------------------ parity ----------

def pcm_to_int(pcm: np.ndarray, bits: int, be: bool) -> np.ndarray:
    nb = bits // 8
    b = pcm.reshape(-1, nb).astype(np.int64)
    if be:
        b = b[:, ::-1]
    v = np.zeros(b.shape[0], dtype=np.int64)
    for k in range(nb):
        v |= b[:, k] << (8 * k)
    sign = 1 << (bits - 1)
    return (v ^ sign) - sign


def lsb_flips(a: np.ndarray, b: np.ndarray, bits: int, be: bool):
    """(#samples that differ, max |difference| in LSB)."""
    d = np.abs(pcm_to_int(a, bits, be) - pcm_to_int(b, bits, be))
    return int(np.count_nonzero(d)), int(d.max(initial=0))


class ParityWindows:
    """Windows of one rank's block, checked against the CPU oracle (the CHECKER, never the thing
    measured): FilterCore.h:57-76 at the seams and edges, ProcessFile.cp:92-101,117 for peak, scale
    and encode."""

    def __init__(self, cfg, blk, total_frames, seed, kernel, rank, W=512, n_random=4):
        import oracle

        self.o = oracle
        oracle.build()
        self.cfg, self.blk, self.total, self.seed = cfg, blk, total_frames, seed
        self.taps = kernel.taps()
        self.H = kernel.half_len
        W = min(W, blk.frames) & ~1
        self.W = W
        rng = np.random.default_rng(1000 * rank + len(cfg["name"]))
        starts = {0, blk.frames - W}                      # both sides of this block's seams / the file edges
        if blk.frames > 3 * W:
            for s in rng.integers(W, blk.frames - 2 * W, n_random):
                starts.add(int(s) & ~1)
        self.starts = sorted(starts)
        self.want = {}                                    # start -> oracle FP64 window [ch][W]
        self.result = {"windows": 0, "samples": 0, "worst_d3": 0.0, "flips": 0, "max_flip_lsb": 0, "ok": True,
                       "pcm_windows": 0}

    def _oracle_window(self, s):
        c, o = self.cfg, self.o
        ch, bits, be, fs = c["channels"], c["bits"], c["be"], c["fs"]
        a = self.blk.start + s                            # absolute frame
        lo, hi = max(0, a - self.H), min(self.total, a + self.W + self.H)     # clipped only at true file ends
        x = o.decode(o.synth_pcm(self.seed, lo, hi - lo, ch, bits, be, fs), hi - lo, ch, bits, be)
        want = np.empty((ch, self.W))
        scale = np.empty((ch, self.W))
        for k in range(ch):
            want[k] = o.fir_hi(x[k], self.taps, a - lo, a - lo + self.W)[a - lo:a - lo + self.W]
            scale[k] = o.fir_abs_scale(x[k], self.taps, a - lo, a - lo + self.W)[a - lo:a - lo + self.W]
        return want, scale

    def check_signal(self, ctx):
        """The parked FP64 signal, D3: |got - want| <= 1e-12 * sum|h x|."""
        ch = self.cfg["channels"]
        for s in self.starts:
            want, scale = self._oracle_window(s)
            self.want[s] = want
            got = ctx.parked_range(s, self.W, ch)
            rel = np.abs(got - want) / np.maximum(scale, 1e-300)
            self.result["worst_d3"] = max(self.result["worst_d3"], float(rel.max()))
            self.result["ok"] &= bool(np.all(rel <= TOL))
            self.result["windows"] += 1
            self.result["samples"] += self.W * ch
        return self

    def check_pcm(self, scale, window_of):
        """Encoded PCM of the same windows against oracle.encode at the (all-reduced) scale;
        window_of(byte_lo, byte_hi) -> uint8 array of this rank's output block."""
        c = self.cfg
        ch, bits, be = c["channels"], c["bits"], c["be"]
        fb = ch * bits // 8
        for s in self.starts:
            want_pcm = self.o.encode(self.want[s], scale, bits, be)
            n, mx = lsb_flips(np.asarray(window_of(s * fb, (s + self.W) * fb)), want_pcm, bits, be)
            self.result["flips"] += n
            self.result["max_flip_lsb"] = max(self.result["max_flip_lsb"], mx)
            # a flip needs y*gain within ~1e-12*scale of a rounding boundary: a couple per window at most
            self.result["ok"] &= mx <= 1 and n <= 4
            self.result["pcm_windows"] += 1
        return self

    def check_taps(self, cfg):
        want = self.o.build_lowcut(cfg["freq"] / cfg["fs"], cfg["slope"] / cfg["fs"])
        ulp = np.abs(self.taps - want) / np.spacing(np.abs(want))
        self.result["taps_differing"] = int(np.count_nonzero(self.taps != want))
        self.result["taps_max_ulp"] = float(ulp.max())
        self.result["ok"] &= bool(ulp.max() <= 1.0) and bool(np.array_equal(self.taps, self.taps[::-1]))
        return self


# ----------------------------------------------------------------- the GPU arm ----------

_REAL_STDOUT = None


def claim_stdout() -> None:
    """stdout carries exactly ONE JSON line.  Libraries print there too (NCCL's version
    banner at communicator start-up, for one), so keep a private handle on the real stdout
    for the result and point fd 1 at stderr for everybody else."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def log(*a) -> None:
    print("[bench]", *a, file=sys.stderr, flush=True)


def peak_hbm() -> float:
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0  # B200_PROFILING.md fallback


def max_over_ranks(world: int, dev, *vals):
    """Max over ranks of each value (timings: the slowest rank counts)."""
    if world == 1:
        return list(vals)
    import torch
    import torch.distributed as dist

    t = torch.tensor(list(vals), dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def gather_rows(world: int, dev, row):
    """Every rank's list of floats -> list of lists, rank by rank (on every rank)."""
    if world == 1:
        return [list(row)]
    import torch
    import torch.distributed as dist

    t = torch.tensor(list(row), dtype=torch.float64, device=dev)
    out = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [[float(v) for v in o] for o in out]


class Env:
    """What every workload of one bench process shares."""

    def __init__(self, a):
        import torch
        import torch.distributed as dist

        from audio_fir_filter_b200 import capi

        self.torch, self.dist, self.capi, self.a = torch, dist, capi, a
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.numa_note = bind_to_gpu_numa_node(self.local) if self.world > 1 else "numa: not bound (single rank)"
        torch.cuda.set_device(self.local)
        self.dev = torch.device(f"cuda:{self.local}")
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            # NCCL's kernels on a high-priority stream: the 8-byte all-reduce of file k's peak must not queue
            # behind the CTAs of file k+1's FIR when two files are in flight (the pipelined arm)
            opts = dist.ProcessGroupNCCL.Options(is_high_priority_stream=True)
            dist.init_process_group("nccl", device_id=self.dev, pg_options=opts)
        self.ctx = capi.Context(self.local)
        self.stream = torch.cuda.Stream(device=self.dev)
        self.ctx.set_stream(self.stream.cuda_stream)
        if a.variant >= 0:
            self.ctx.set_variant(a.variant)
        if a.codec_tile or a.codec_threads or a.codec_carveout != 50:
            self.ctx.set_codec_geometry(a.codec_tile, a.codec_threads or 256, a.codec_carveout)
        self._ctx2 = None
        # measured FP64 ceilings of this GPU, before the run (register-resident probes), with the
        # clocks the GPU ran at while they were measured (they are the roofline's denominator)
        sampler = ClockSampler(self.local) if self.rank == 0 else None
        t0 = time.perf_counter()
        self.dfma_peak = self.ctx.fp64_peak(0, 0.4)
        self.dmma_peak = self.ctx.fp64_peak(1, 0.4)
        self.probe_clocks = sampler.stop(t0, time.perf_counter()) if sampler else None

    @property
    def ctx2(self):
        """A second context on the same GPU (the pipelined end-to-end arm)."""
        if self._ctx2 is None:
            self._ctx2 = self.capi.Context(self.local)
            self._stream2 = self.torch.cuda.Stream(device=self.dev)
            self._ctx2.set_stream(self._stream2.cuda_stream)
            if self.a.variant >= 0:
                self._ctx2.set_variant(self.a.variant)
        return self._ctx2

    def barrier(self):
        self.torch.cuda.synchronize(self.dev)
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, *vals):
        return max_over_ranks(self.world, self.dev, *vals)

    def gather_rows(self, row):
        return gather_rows(self.world, self.dev, row)

    def close(self):
        if self._ctx2 is not None:
            self._ctx2.close()
        self.ctx.close()


def run_workload(env: Env, cfg_id: int, *, steps: int, warmup: int, mode: str = "block", world: int | None = None,
                 strong: bool = False, e2e_steps: int | None = None, e2e_warmup: int | None = None,
                 pipelined: bool = False, sample_clocks: bool = False, parity: bool = True,
                 copy_ceiling: bool = False) -> dict:
    """One workload on `world` ranks (this process = one of them): device-resident arm, parity,
    end-to-end arm(s), roofline.  Returns the record (complete on every rank)."""
    from audio_fir_filter_b200.dist import allreduce_max_peak
    from audio_fir_filter_b200.process import plan_blocks, scale_for_peak

    torch, capi, a = env.torch, env.capi, env.a
    cfg = CONFIGS[cfg_id]
    world = env.world if world is None else world
    rank = env.rank if world > 1 else 0
    ctx, dev, stream = env.ctx, env.dev, env.stream
    inplace = cfg_id == 6            # the encoded PCM overwrites the input block: 180 GB would not hold both
    fs, ch, bits, be = cfg["fs"], cfg["channels"], cfg["bits"], cfg["be"]
    fb = ch * bits // 8
    taps = kernel_order(cfg["slope"] / fs) + 1
    H = (taps - 1) // 2
    if world > 1 and mode == "block":
        total_frames = cfg["frames"] if strong else cfg["frames"] * world
        blk = plan_blocks(total_frames, world, H)[rank]
    else:
        total_frames = cfg["frames"]
        blk = plan_blocks(total_frames, 1, H)[0]
    first = blk.start - blk.halo_left
    n_in = blk.total_frames
    in_bytes, out_bytes = n_in * fb, blk.frames * fb
    collective = world > 1 and mode == "block"
    seed = SEED if mode == "block" else SEED + rank

    kernel = ctx.build_kernel(cfg["freq"] / fs, cfg["slope"] / fs)
    assert kernel.num_taps == taps
    d_in = torch.empty(in_bytes, dtype=torch.uint8, device=dev)
    d_out = d_in[:out_bytes] if inplace else torch.empty(out_bytes, dtype=torch.uint8, device=dev)

    def synth():
        ctx.synth_pcm_dev(seed, first, n_in, ch, bits, be, fs, 1.0, d_in)
        ctx.synchronize()

    synth()
    # a pass that is timed once (config 3 in full) must not time cudaMalloc of its 30+ GB of buffers
    ctx.reserve(kernel, blk.frames, ch, bits, be, blk.halo_left, blk.halo_right,
                host_path=not (a.no_e2e or e2e_steps == 0 or inplace))

    def barrier():
        if world > 1:
            env.barrier()
        else:
            torch.cuda.synchronize(dev)

    def reduce_peak(c=ctx) -> float:
        return allreduce_max_peak(c) if collective else c.peak()

    def step_dev() -> float:
        ctx.apply_dev(kernel, d_in, blk.frames, ch, bits, be, blk.halo_left, blk.halo_right)
        pk = reduce_peak()
        ctx.encode_dev(scale_for_peak(pk, cfg["normalize"]), d_out)
        return pk

    # ---- kernel-resident arm (`value`) -------------------------------------------------
    sampler = ClockSampler(env.local) if (sample_clocks and env.rank == 0) else None
    for _ in range(warmup):
        step_dev()
        if inplace:
            synth()                  # the pass consumed its input
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fir_ms, dec_ms, enc_ms = [], [], []
    launches = 0
    e0.record(stream)
    t0 = time.perf_counter()
    for _ in range(steps):
        pk = step_dev()
        t = ctx.last_timing()            # waits for the step; per-kernel CUDA-event spans
        fir_ms.append(t["fir_ms"])
        dec_ms.append(t["decode_ms"])
        enc_ms.append(t["encode_ms"])
        launches += int(t["fir_launches"] + t["other_launches"])
    e1.record(stream)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = None
    if sampler:
        # a timed region shorter than a few sampler periods (config 1: ~3 ms a step) cannot be
        # sampled: repeat the identical step untimed, right away, until the sampler has seen load
        note = None
        t_rep = t_end = time.perf_counter()
        while sampler.count_since(t0) < 4 and time.perf_counter() - t_rep < 3.0 and world == 1:
            step_dev()
            t_end = time.perf_counter()
            note = "timed region shorter than the sampling period: sampled over an identical untimed repeat right after it"
        clocks = sampler.stop(t0, t_end)
        if note:
            clocks["note"] = note
    dev_ms, wall_ms = env.max_over_ranks(e0.elapsed_time(e1), wall_ms) if world > 1 else (e0.elapsed_time(e1), wall_ms)
    ms_per_step = dev_ms / steps
    out_samples_step = total_frames * ch if (mode == "block" or world == 1) else cfg["frames"] * ch * world
    value = out_samples_step / (ms_per_step * 1e-3) / 1e6
    scale = scale_for_peak(pk, cfg["normalize"])

    # ---- parity of what the device arm just produced -------------------------------------
    par = None
    pw = None
    if parity:
        tp = time.perf_counter()
        pw = ParityWindows(cfg, blk, total_frames, seed, kernel, rank)
        pw.check_taps(cfg).check_signal(ctx)
        pw.check_pcm(scale, lambda lo, hi: d_out[lo:hi].cpu().numpy())
        # this block's own peak from the stand-alone warp-reduced kernel.  Without a collective the
        # device scalar still holds the value fused into the FIR epilogue: the two must be equal.  After
        # the all-reduce it holds the GLOBAL peak; then max over ranks of the stand-alone block peaks must
        # equal it (checked once the rows are gathered) -- fused epilogues and NCCL MAX verified together.
        local_peak = ctx.peak_recompute()
        if not collective:
            pw.result["ok"] &= ctx.peak() == local_peak
        pw.result["peak_local"] = local_peak
        pw.result["device_check_s"] = time.perf_counter() - tp

    # ---- end-to-end arm through the host-buffer C-ABI ------------------------------------
    e2e = None
    if not a.no_e2e and e2e_steps != 0 and not inplace:
        ks = steps if e2e_steps is None else e2e_steps
        kw = warmup if e2e_warmup is None else e2e_warmup
        h_in = capi.PinnedBuffer(in_bytes)
        h_out = capi.PinnedBuffer(out_bytes)
        torch.from_numpy(h_in.array).copy_(d_in)          # cudaMemcpy straight into the pinned buffer
        one_call = world == 1 or mode == "batch"

        def step_host():
            if one_call:
                # the one-call entry point a host makes per file (fir_gpu_process)
                p, _ = ctx.process(kernel, h_in.array, blk.frames, ch, bits, be, cfg["normalize"], h_out.array)
                return p
            ctx.apply(kernel, h_in.array, blk.frames, ch, bits, be, blk.halo_left, blk.halo_right)
            p = reduce_peak()
            ctx.encode(scale_for_peak(p, cfg["normalize"]), h_out.array)   # synchronous: D2H done on return
            return p

        for _ in range(kw):
            step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(ks):
            pk_host = step_host()
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        (e2e_ms,) = env.max_over_ranks(e2e_ms) if world > 1 else (e2e_ms,)
        e2e_ms /= ks
        th = ctx.last_timing()
        same = bool(np.array_equal(h_out.array, d_out.cpu().numpy())) and pk_host == pk
        e2e = {"value": out_samples_step / (e2e_ms * 1e-3) / 1e6, "unit": "MSamples/s", "ms_per_step": e2e_ms,
               "steps": ks, "warmup": kw,
               "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes + 8,
               "h2d_ms": th["h2d_ms"], "d2h_ms": th["d2h_ms"],
               "d2h_gbs_in_step": out_bytes / (th["d2h_ms"] * 1e-3) / 1e9 if th["d2h_ms"] > 0 else None,
               "entry_point": "fir_gpu_process" if one_call else "fir_gpu_apply + NCCL all-reduce MAX + fir_gpu_encode",
               "host_buffers": "pinned (fir_gpu_host_alloc); " + env.numa_note,
               "matches_device_arm": same}
        if pw is not None:
            pw.check_pcm(scale, lambda lo, hi: h_out.array[lo:hi])            # the host arm's bytes vs the oracle
            pw.result["ok"] &= same

        # -- what plain pinned copies of the same size reach on this host, all ranks at once --
        if copy_ceiling:
            cc = {}
            for name, direction, buf, n in (("d2h", 1, h_out, out_bytes), ("h2d", 0, h_in, in_bytes)):
                best = None
                for _ in range(3):
                    barrier()
                    ms = ctx.copy_probe(buf.array, n, direction)
                    (ms,) = env.max_over_ranks(ms) if world > 1 else (ms,)
                    best = ms if best is None else min(best, ms)
                cc[f"{name}_ms"] = best
                cc[f"{name}_gbs_per_rank"] = n / (best * 1e-3) / 1e9
                cc[f"{name}_gbs_aggregate"] = world * n / (best * 1e-3) / 1e9
            # the same download into a buffer pinned by somebody else's allocator (torch: cudaHostAlloc
            # with default flags, not cudaHostAllocPortable): is the rate a property of OUR buffers?
            tp = torch.empty(out_bytes, dtype=torch.uint8, pin_memory=True)
            best = None
            for _ in range(3):
                barrier()
                ms = ctx.copy_probe(tp.data_ptr(), out_bytes, 1)
                (ms,) = env.max_over_ranks(ms) if world > 1 else (ms,)
                best = ms if best is None else min(best, ms)
            cc["d2h_gbs_per_rank_torch_pinned_buffer"] = out_bytes / (best * 1e-3) / 1e9
            del tp
            if world > 1:                                                   # one rank alone, the others idle
                barrier()
                alone = ctx.copy_probe(h_out.array, out_bytes, 1) if rank == 0 else 0.0
                barrier()
                (alone,) = env.max_over_ranks(alone)
                cc["d2h_gbs_rank0_alone"] = out_bytes / (alone * 1e-3) / 1e9
            cc["note"] = ("one cudaMemcpyAsync of the step's bytes per rank between the same pinned buffers and HBM, "
                          "all ranks at the same moment (max over ranks, best of 3)")
            e2e["copy_ceiling"] = cc

        # -- pipelined: two contexts per rank alternate files, so that file k's encode + download
        #    run under file k+1's FIR (what GpuPool::LANES does in the C++ host).  Every step's
        #    copies are still inside the timed region. --
        if pipelined and ks >= 2:
            ctxs = [ctx, env.ctx2]
            h_outs = [h_out, capi.PinnedBuffer(out_bytes)]

            def start(i):
                ctxs[i & 1].apply(kernel, h_in.array, blk.frames, ch, bits, be, blk.halo_left, blk.halo_right)

            def run_pipe(n):
                # two files in flight: while file i's FIR runs, file i+1 is already uploaded and queued
                # behind it; file i's encode + download and file i+2's upload then run under FIR i+1
                start(0)
                if n > 1:
                    start(1)
                p = 0.0
                for i in range(n):
                    cur = ctxs[i & 1]
                    p = reduce_peak(cur)                                     # waits for FIR i
                    cur.encode(scale_for_peak(p, cfg["normalize"]), h_outs[i & 1].array)   # under FIR i+1
                    if i + 2 < n:
                        start(i + 2)                                         # its upload runs under FIR i+1 as well
                return p

            run_pipe(max(2, kw))
            barrier()
            t0 = time.perf_counter()
            pk_pipe = run_pipe(ks)
            barrier()
            pipe_ms = (time.perf_counter() - t0) * 1e3
            (pipe_ms,) = env.max_over_ranks(pipe_ms) if world > 1 else (pipe_ms,)
            pipe_ms /= ks
            ok = all(bool(np.array_equal(h.array, h_out.array)) for h in h_outs) and pk_pipe == pk
            e2e["pipelined"] = {"value": out_samples_step / (pipe_ms * 1e-3) / 1e6, "unit": "MSamples/s",
                                "ms_per_step": pipe_ms, "contexts_per_gpu": 2, "matches_serial_arm": ok,
                                "what": "same steps with two files in flight per rank (two contexts alternating): encode + "
                                        "download of file k and the upload of file k+2 run under the FIR of file k+1; all "
                                        "copies inside the timed region"}
            if pw is not None:
                pw.result["ok"] &= ok
            h_outs[1].free()
        h_in.free()
        h_out.free()

    # ---- gather the parity blocks of all ranks ----------------------------------------------
    if pw is not None:
        r = pw.result
        rows = env.gather_rows([r["windows"], r["samples"], r["worst_d3"], r["flips"], r["max_flip_lsb"],
                                1.0 if r["ok"] else 0.0, r["peak_local"], r["pcm_windows"], r["taps_differing"],
                                r["taps_max_ulp"]]) if world > 1 else [[r["windows"], r["samples"], r["worst_d3"],
                                                                        r["flips"], r["max_flip_lsb"],
                                                                        1.0 if r["ok"] else 0.0, r["peak_local"],
                                                                        r["pcm_windows"], r["taps_differing"],
                                                                        r["taps_max_ulp"]]]
        peaks = [x[6] for x in rows]
        peak_ok = (max(peaks) == pk) if (collective or world == 1) else True   # ProcessFile.cp:92-96 over ALL blocks
        par = {
            "ok": all(x[5] == 1.0 for x in rows) and peak_ok,
            "ranks": len(rows), "windows": int(sum(x[0] for x in rows)), "samples": int(sum(x[1] for x in rows)),
            "window_frames": pw.W, "worst_d3": max(x[2] for x in rows), "tolerance_d3": TOL,
            "pcm_windows": int(sum(x[7] for x in rows)), "flips": int(sum(x[3] for x in rows)),
            "max_flip_lsb": int(max(x[4] for x in rows)),
            "taps_differing_from_oracle": int(rows[0][8]), "taps_max_ulp": rows[0][9],
            "global_peak_is_max_of_block_peaks": peak_ok, "peak": pk, "scale": scale,
            "per_rank": [{"windows": int(x[0]), "worst_d3": x[2], "flips": int(x[3]), "ok": x[5] == 1.0,
                          "block_peak": x[6]} for x in rows],
            "what": "per rank: first and last window of its block (= both sides of every seam, true file edges on the "
                    "outer ranks) + random interior windows; parked FP64 signal vs long-double oracle on the same "
                    "synthetic PCM (D3), encoded PCM of the device arm and of the host arm vs oracle encode at the "
                    "all-reduced scale, fused peak == stand-alone peak kernel, taps vs oracle (<= 1 ulp)",
        }

    # ---- roofline of the dominant kernel (FIR) -------------------------------------------
    flop_launch = algorithmic_flop(blk.frames, ch, taps, blk.halo_left, blk.halo_right)
    fir_avg_ms = sum(fir_ms) / len(fir_ms)
    achieved = flop_launch / (fir_avg_ms * 1e-3) / 1e12
    vname = capi.variant_names()[a.variant if a.variant >= 0 else 0]
    is_dmma = vname.startswith("dmma")
    fp64_peak = env.dmma_peak if is_dmma else env.dfma_peak
    traffic, traffic_source = None, None
    tpath = os.path.join(ROOT, "profiles", "fir_traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = tj.get(f"cfg{cfg_id}")
            traffic_source = tj.get("source") if traffic is not None else None
        except Exception:
            traffic = None
    roofline = {
        "kernel": "fir_dmma_kernel" if is_dmma else "fir_fp64_kernel",
        # the FP64 tensor pipe (DMMA.8x8x4) for the default kernel, the FP64 FMA pipe for the dfma_* variants
        "bound": "tensor" if is_dmma else "fp64_fma",
        "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
        "frac": achieved / fp64_peak, "traffic": traffic,
        "traffic_source": (traffic_source or "none for this config") +
                          " -- a constant from that ncu capture, NOT measured in this run",
        "algorithmic_bytes_per_launch": 16.0 * blk.frames * ch,
        "peak_source": "measured live on this GPU before the run: register-resident "
                       + ("DMMA m8n8k4 f64 probe" if is_dmma else "DFMA probe")
                       + " (fir_gpu_fp64_peak); MEASURED_PEAKS.json has no FP64 figure (its bf16 number is for "
                         "tcgen05, which has no FP64 kind)",
        "dfma_probe": env.dfma_peak, "probe_clocks": env.probe_clocks,
        "peak_nominal": 148 * 64 * 2 * 1.965e9 / 1e12, "dmma_probe": env.dmma_peak,
        "frac_of_nominal": achieved / (148 * 64 * 2 * 1.965e9 / 1e12),
        "flop_per_launch": flop_launch, "fir_ms_per_launch": fir_avg_ms,
        "fir_share_of_step": fir_avg_ms / ms_per_step, "fir_variant": vname,
        "codec_hbm": {
            "decode_gbs": (n_in * fb + (blk.frames + 2 * H) * ch * 8) / (sum(dec_ms) / len(dec_ms) * 1e-3) / 1e9,
            "encode_gbs": (blk.frames * ch * 8 + out_bytes) / (sum(enc_ms) / len(enc_ms) * 1e-3) / 1e9,
            "decode_ms": sum(dec_ms) / len(dec_ms), "encode_ms": sum(enc_ms) / len(enc_ms),
            "peak_gbs": peak_hbm(),
        },
    }
    roofline["codec_hbm"]["decode_frac"] = roofline["codec_hbm"]["decode_gbs"] / roofline["codec_hbm"]["peak_gbs"]
    roofline["codec_hbm"]["encode_frac"] = roofline["codec_hbm"]["encode_gbs"] / roofline["codec_hbm"]["peak_gbs"]

    kernel.free()
    del d_in, d_out
    torch.cuda.empty_cache()
    return {
        "value": value, "ms_per_step": ms_per_step, "steps": steps, "warmup": warmup, "n_gpus": world,
        "config": config_dict(cfg, world, mode, strong),
        "fp64_tflops": flop_launch * world / (ms_per_step * 1e-3) / 1e12,
        "host_wall_ms_per_step": wall_ms / steps, "gpu_launches": launches, "clocks": clocks, "e2e": e2e,
        "roofline": roofline, "parity": par, "peak_value": pk,
    }


def also_record(r: dict) -> dict:
    """The compact per-config record of the `also` block."""
    rf, e = r["roofline"], r["e2e"]
    return {
        "workload": r["config"]["workload"], "taps": r["config"]["taps"], "n_gpus": r["n_gpus"],
        "steps": r["steps"], "warmup": r["warmup"],
        "value": r["value"], "unit": "MSamples/s", "ms_per_step": r["ms_per_step"],
        "fir_ms": rf["fir_ms_per_launch"], "tflops": rf["achieved"], "frac": rf["frac"],
        "frac_of_nominal": rf["frac_of_nominal"], "fp64_tflops_all_gpus": r["fp64_tflops"],
        "kernel_only_ms": rf["fir_ms_per_launch"] + rf["codec_hbm"]["decode_ms"] + rf["codec_hbm"]["encode_ms"],
        "e2e_ms": e["ms_per_step"] if e else None, "e2e_value": e["value"] if e else None,
        "e2e_pipelined_ms": e["pipelined"]["ms_per_step"] if e and "pipelined" in e else None,
        "h2d_ms": e["h2d_ms"] if e else None, "d2h_ms": e["d2h_ms"] if e else None,
        "codec_hbm": rf["codec_hbm"], "gpu_launches": r["gpu_launches"],
        "parity": {k: v for k, v in r["parity"].items() if k not in ("per_rank", "what")} if r["parity"] else None,
    }


# ------------------------------------------ the shipped CLI on N GPUs vs one GPU ----------

def wav_file_bytes(pcm: bytes, channels: int, bits: int, rate: int) -> bytes:
    """A minimal RIFF/WAVE with a foreign chunk either side of `data` (metadata must survive)."""
    nb = bits // 8

    def chunk(cid, data):
        return cid + struct.pack("<I", len(data)) + data + (b"\0" if len(data) & 1 else b"")

    body = chunk(b"fmt ", struct.pack("<HHIIHH", 1, channels, rate, rate * channels * nb, channels * nb, bits))
    body += chunk(b"bext", b"bench.py\0" * 5) + chunk(b"data", pcm) + chunk(b"LIST", b"INFOINAM\x06\0\0\0bench\0")
    return b"RIFF" + struct.pack("<I", 4 + len(body)) + b"WAVE" + body


def aiff_file_bytes(pcm: bytes, channels: int, bits: int, rate: int) -> bytes:
    """A minimal FORM/AIFF (COMM + SSND) around big-endian PCM, with a foreign chunk either side."""
    def chunk(cid, data):
        return cid + struct.pack(">I", len(data)) + data + (b"\0" if len(data) & 1 else b"")

    frames = len(pcm) // (channels * bits // 8)
    m, e = math.frexp(float(rate))                          # 80-bit extended sample rate
    ext = struct.pack(">HQ", e - 1 + 16383, int(m * (1 << 64)))
    body = chunk(b"NAME", b"bench.py") + chunk(b"COMM", struct.pack(">hIh", channels, frames, bits) + ext)
    body += chunk(b"SSND", struct.pack(">II", 0, 0) + pcm) + chunk(b"ANNO", b"after the samples")
    return b"FORM" + struct.pack(">I", 4 + len(body)) + b"AIFF" + body


def cli_single_file(env: Env, cfg: dict) -> dict:
    """What a `lowcut` user waits for on the headline workload: the shipped C++ host on a tmpfs file
    of the config's shape, wall clock, with the start-up / filter / write / exit break-down of its -v
    time stamps -- and its output compared, byte for byte, with what this process's own context makes
    of the same payload (which the parity block has checked against the oracle)."""
    import re

    torch, ctx = env.torch, env.ctx
    lowcut = os.path.join(ROOT, "host", "lowcut")
    if not os.path.exists(lowcut):
        return {"ok": False, "error": "host/lowcut is not built"}
    fs, ch, bits, be, frames = cfg["fs"], cfg["channels"], cfg["bits"], cfg["be"], cfg["frames"]
    fb = ch * bits // 8
    d = torch.empty(frames * fb, dtype=torch.uint8, device=env.dev)
    ctx.synth_pcm_dev(SEED, 0, frames, ch, bits, be, fs, 1.0, d)
    ctx.synchronize()
    pcm = d.cpu().numpy()
    del d
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    tmp = tempfile.mkdtemp(prefix="lowcut_bench_", dir=base)
    try:
        ext = ".aif" if be else ".wav"
        src, out = os.path.join(tmp, "in" + ext), os.path.join(tmp, "out" + ext)
        data = aiff_file_bytes(pcm.tobytes(), ch, bits, fs) if be else wav_file_bytes(pcm.tobytes(), ch, bits, fs)
        open(src, "wb").write(data)
        args = [lowcut, "-O", "-v", "-f", str(cfg["freq"]), "-s", str(cfg["slope"])] + (["-n"] if cfg["normalize"] else [])
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            r = subprocess.run(args + [src, out], capture_output=True, text=True, timeout=600)
            dt = time.perf_counter() - t0
            if r.returncode != 0:
                return {"ok": False, "error": f"lowcut exit {r.returncode}: {r.stderr[-300:]}"}
            if best is None or dt < best[0]:
                best = (dt, r.stdout)
        wall, text = best

        def stamp(pat):
            m = re.search(pat, text)
            return float(m.group(1)) if m else None

        ready = stamp(r"\[\s*([\d.]+) s since start\] \d+ GPU context")
        done = stamp(r"\[\s*([\d.]+) s since start\] done")
        filtered = stamp(r"\[\s*([\d.]+) s\] filtered")
        written = stamp(r"\[\s*([\d.]+) s\] encoded and written")
        off = data.index(b"SSND") + 16 if be else data.index(b"data") + 8
        got = open(out, "rb").read()
        k = ctx.build_kernel(cfg["freq"] / fs, cfg["slope"] / fs)
        want = np.empty(pcm.size, dtype=np.uint8)
        ctx.process(k, pcm, frames, ch, bits, be, cfg["normalize"], want)
        k.free()
        same = got[off:off + pcm.size] == want.tobytes()
        meta = got[:off] == data[:off] and got[off + pcm.size:] == data[off + pcm.size:]
        return {"workload": cfg["name"], "file": f"{len(data) / 1e6:.1f} MB on " + ("tmpfs" if base else "a tmp dir"),
                "wall_s": wall, "msamples_per_s_wall": frames * ch / wall / 1e6,
                "start_up_s": ready, "read_upload_filter_s": filtered,
                "encode_download_write_s": (written - filtered) if written is not None and filtered is not None else None,
                "exit_s": (wall - done) if done is not None else None,
                "payload_equals_library": bool(same), "metadata_identical": bool(meta), "ok": bool(same and meta),
                "what": "best of 3 runs of host/lowcut -v; start_up = exec + CUDA init + context; exit = the driver "
                        "releasing the context after the output is complete"}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def cli_block_identity(env: Env, n_gpus: int) -> dict:
    """host/lowcut -g N against -g 1 on one tmpfs file long enough for sample-block mode: the two
    outputs must be byte-identical (FilterCore.h:57-76 at the seams, ProcessFile.cp:92-101: one
    peak over all blocks -- the file is filtered with -n).  The payload is also run through this
    process's own context and its seam windows are checked against the oracle."""
    import oracle
    from audio_fir_filter_b200.process import plan_blocks

    torch, ctx = env.torch, env.ctx
    lowcut = os.path.join(ROOT, "host", "lowcut")
    if not os.path.exists(lowcut):
        return {"ok": False, "error": "host/lowcut is not built"}
    fs, ch, bits, be = 48000, 2, 24, False
    frames = max(2_880_000, n_gpus * ((1 << 18) + 4096))
    fb = ch * bits // 8
    d = torch.empty(frames * fb, dtype=torch.uint8, device=env.dev)
    ctx.synth_pcm_dev(SEED + 7, 0, frames, ch, bits, be, fs, 1.0, d)
    ctx.synchronize()
    pcm = d.cpu().numpy()
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    tmp = tempfile.mkdtemp(prefix="lowcut_bench_", dir=base)
    res = {"file": f"{frames} frames x {ch} ch {bits}-bit LE WAV, -f 20 -s 20 -n, tmpfs" if base else "tmp dir"}
    try:
        src = os.path.join(tmp, "in.wav")
        file_bytes = wav_file_bytes(pcm.tobytes(), ch, bits, fs)
        open(src, "wb").write(file_bytes)
        outs, walls = {}, {}
        for g in (1, n_gpus):
            out = os.path.join(tmp, f"out_g{g}.wav")
            t0 = time.perf_counter()
            # LOWCUT_NCCL=1: take the peak over the blocks with the NCCL all-reduce even though this file is too
            # short for the communicator start-up to hide (the host would default to a host max of the scalars)
            r = subprocess.run([lowcut, "-g", str(g), "-n", "-v", "-f", "20", "-s", "20", src, out], capture_output=True,
                               text=True, timeout=600, env=dict(os.environ, LOWCUT_NCCL="1"))
            walls[g] = time.perf_counter() - t0
            if r.returncode != 0:
                return {**res, "ok": False, "error": f"lowcut -g {g} exit {r.returncode}: {r.stderr[-300:]}"}
            outs[g] = open(out, "rb").read()
            if g == n_gpus:                                   # -v prints one "device time" line per block
                res["blocks_used"] = sum(1 for l in r.stdout.splitlines() if "device time" in l)
                res["peak_exchange"] = next((l.split("peak exchange:")[1].strip() for l in r.stdout.splitlines()
                                             if "peak exchange:" in l), None)
        off = file_bytes.index(b"data") + 8
        n = frames * fb
        a1, aN = outs[1], outs[n_gpus]
        res["byte_identical"] = a1 == aN
        res["metadata_identical"] = aN[:off] == file_bytes[:off] and aN[off + n:] == file_bytes[off + n:]
        res["wall_s_g1"], res[f"wall_s_g{n_gpus}"] = walls[1], walls[n_gpus]
        # the same payload through this process's context: same bytes as the CLI, and its seam
        # windows against the oracle at the scale of the whole file's peak
        k = ctx.build_kernel(20.0 / fs, 20.0 / fs)
        h_out = np.empty(n, dtype=np.uint8)
        pk, sc = ctx.process(k, pcm, frames, ch, bits, be, True, h_out)
        res["cli_equals_library"] = bytes(aN[off:off + n]) == h_out.tobytes()
        taps, H, W = k.taps(), k.half_len, 512
        worst, flips, mx = 0.0, 0, 0
        payload = np.frombuffer(aN, dtype=np.uint8, count=n, offset=off)
        for b in plan_blocks(frames, n_gpus, H):
            for s in {b.start, b.start + b.frames - W}:
                lo, hi = max(0, s - H), min(frames, s + W + H)
                x = oracle.decode(pcm[lo * fb:hi * fb], hi - lo, ch, bits, be)
                want = np.stack([oracle.fir_hi(x[c], taps, s - lo, s - lo + W)[s - lo:s - lo + W] for c in range(ch)])
                nf, m = lsb_flips(payload[s * fb:(s + W) * fb], oracle.encode(want, sc, bits, be), bits, be)
                flips += nf
                mx = max(mx, m)
        k.free()
        res.update({"seam_windows": 2 * n_gpus, "seam_flips": flips, "seam_max_flip_lsb": mx, "peak": pk})
        res["ok"] = bool(res["byte_identical"] and res["metadata_identical"] and res["cli_equals_library"] and mx <= 1
                         and flips <= 4 * n_gpus and res.get("blocks_used", 0) == n_gpus
                         and "ncclAllReduce" in (res.get("peak_exchange") or ""))
        return res
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


# --------------------------------------------------------------------- main ----------

def main() -> int:
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="block", choices=["block", "batch"],
                    help="N>1: one long file in sample blocks with halo + peak all-reduce, or a file per rank")
    ap.add_argument("--also", default="auto", choices=["auto", "none", "small", "all"],
                    help="extra one-pass records of the other BASELINE configs (auto: all at N=1 with the default "
                         "config, config 3 block-split at N=8, none otherwise)")
    ap.add_argument("--variant", type=int, default=-1, help="FIR kernel variant (experiments)")
    ap.add_argument("--codec-tile", type=int, default=0, help="codec tile bytes (experiments)")
    ap.add_argument("--codec-threads", type=int, default=0, help="codec threads per CTA (experiments)")
    ap.add_argument("--codec-carveout", type=int, default=50, help="codec shared-memory carveout percent (experiments)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle windows (profiling runs)")
    ap.add_argument("--no-cli", action="store_true", help="skip the host/lowcut runs (N=1: timing + identity on the "
                                                          "workload's file; N>1: -g N vs -g 1)")
    a = ap.parse_args()
    if a.warmup < 3 and a.config != 6:
        a.warmup = 3 if a.impl == "ours" else a.warmup
    cfg = CONFIGS[a.config]
    if a.config == 6:
        a.steps, a.warmup, a.no_e2e, a.no_cpu = 1, 1, True, True
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if a.impl == "reference":
        return reference_arm(a, cfg, rank)

    import torch

    if not torch.cuda.is_available():
        print("bench.py: no CUDA device; this framework has no CPU path", file=sys.stderr)
        return 2
    if a.config == 6 and world != 8:
        print("bench.py: --config 6 (config 5 in full) needs --gpus 8", file=sys.stderr)
        return 2
    env = Env(a)
    parity = not a.no_parity

    t_all = time.perf_counter()
    head = run_workload(env, a.config, steps=a.steps, warmup=a.warmup, mode=a.mode, pipelined=True,
                        sample_clocks=True, parity=parity, copy_ceiling=True)
    log(f"headline done in {time.perf_counter() - t_all:.1f} s: {head['value']:.1f} MSamples/s, parity "
        f"{head['parity']['ok'] if head['parity'] else 'skipped'}")

    # ---- one record per other BASELINE config ------------------------------------------------
    also, also_mode = {}, a.also
    if also_mode == "auto":
        also_mode = "all" if (a.config == 2 and a.mode == "block" and world in (1, 8)) else "none"
    if also_mode != "none" and world == 1:
        plan = [(1, dict(steps=10, warmup=3, pipelined=True)), (4, dict(steps=5, warmup=2, pipelined=True)),
                (5, dict(steps=1, warmup=1, e2e_steps=1, e2e_warmup=0))]
        if also_mode == "all":
            plan.append((3, dict(steps=1, warmup=0, e2e_steps=1, e2e_warmup=0)))   # config 3 IN FULL, ~29 s a pass
        for cid, kw in plan:
            if cid == a.config:
                continue
            t0 = time.perf_counter()
            if cid == 3 and _mem_available_gb() < 40:
                kw = dict(kw, e2e_steps=0)          # two 8.3 GB pinned buffers would not fit this host
            r = run_workload(env, cid, parity=parity, **kw)
            also[f"cfg{cid}"] = also_record(r)
            also[f"cfg{cid}"]["bench_seconds"] = time.perf_counter() - t0
            log(f"also cfg{cid}: {r['ms_per_step']:.1f} ms/step, {r['roofline']['achieved']:.2f} TFLOP/s, parity "
                f"{r['parity']['ok'] if r['parity'] else 'skipped'} ({time.perf_counter() - t0:.1f} s)")
    if also_mode != "none" and world > 1 and a.mode == "block":   # auto: at N=8 only; --also all: at any N>1
        t0 = time.perf_counter()
        r = run_workload(env, 3, steps=1, warmup=1, strong=True, e2e_steps=1, e2e_warmup=0, parity=parity)
        also["cfg3_block_split"] = also_record(r)
        also["cfg3_block_split"]["bench_seconds"] = time.perf_counter() - t0

    # ---- the shipped C++ host on N GPUs vs one (rank 0; the others wait on the store) -----------
    cli = None
    if world > 1 and not a.no_cli and a.mode == "block":
        store = env.dist.distributed_c10d._get_default_store()
        if rank == 0:
            try:
                cli = cli_block_identity(env, world)
            except Exception as e:  # noqa: BLE001
                cli = {"ok": False, "error": f"{type(e).__name__}: {e}"}
            store.set("bench_cli_done", "1")
        else:
            store.wait(["bench_cli_done"])
        env.barrier()
    if world == 1 and not a.no_cli:
        try:
            cli = cli_single_file(env, cfg)
        except Exception as e:  # noqa: BLE001
            cli = {"ok": False, "error": f"{type(e).__name__}: {e}"}

    line = None
    ok = True
    if rank == 0:
        line = {
            "metric": METRIC, "value": head["value"], "unit": "MSamples/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": head["config"],
            "fp64_tflops": head["fp64_tflops"], "host_wall_ms_per_step": head["host_wall_ms_per_step"],
            "gpu_launches": head["gpu_launches"], "clocks": head["clocks"], "e2e": head["e2e"],
            "roofline": head["roofline"], "parity": head["parity"], "peak_value": head["peak_value"],
        }
        if also:
            line["also"] = also
        if cli is not None:
            line["cli_block_mode" if world > 1 else "cli"] = cli
        ok = (head["parity"] is None or head["parity"]["ok"]) and all(
            v["parity"] is None or v["parity"]["ok"] for v in also.values()) and (cli is None or cli["ok"])
        line["bench_seconds_gpu_arm"] = time.perf_counter() - t_all

    if rank == 0 and not a.no_cpu and world == 1:   # the CPU leg is reported at N=1 only
        rows = cpu_rows(cfg, a.cpu_seconds)
        line["cpu_baseline"] = {"value": rows[0]["value"], "unit": "MSamples/s", "kind": rows[0]["kind"],
                                "cores": rows[0]["threads"], "sample": rows[0]["sample"], "rows": rows,
                                "times": "PCM in -> PCM out: decode + FIR + peak + normalise + encode (the FIR is "
                                         f"{rows[0]['fir_share'] * 100:.1f} % of it)"}
    if world > 1:
        env.dist.barrier()
        env.dist.destroy_process_group()
    if rank == 0:
        emit(line)
        if not ok:
            print("bench.py: PARITY FAILURE -- see the parity / also / cli_block_mode blocks of the line", file=sys.stderr)
    env.close()
    return 0 if ok else 3


def _mem_available_gb() -> float:
    try:
        for l in open("/proc/meminfo"):
            if l.startswith("MemAvailable:"):
                return float(l.split()[1]) / 1e6
    except OSError:
        pass
    return 0.0


def reference_arm(a, cfg, rank: int) -> int:
    """bench.py --impl reference: the reference's CPU implementation of the path on the host
    cores, rank 0 only.  Each step is a bounded sample of the workload, PCM in -> PCM out
    (decode, FIR, peak, normalise, encode all timed), with every host thread; the line also
    carries the reference's default thread count (floor(0.7*cores), main.cp:75) and, where it
    fits in half a minute, the WHOLE file timed once."""
    if rank != 0:
        return 0
    import oracle

    oracle.build()
    cores = len(os.sched_getaffinity(0))
    # size one step for ~ (cpu-seconds / steps), at least 2 s
    per_step = max(2.0, a.cpu_seconds * 2.0 / max(1, a.steps))
    for _ in range(max(0, min(a.warmup, 1))):
        cpu_reference_rate(cfg, 1.0, cores)
    vals, secs, desc = [], [], None
    for _ in range(a.steps):
        v, desc, t = cpu_reference_rate(cfg, per_step, cores)
        vals.append(v)
        secs.append(t)
    value = statistics.median(vals)
    rows = [{"threads": cores, "value": value, "unit": "MSamples/s", "row": "all cores (the line's value)",
             "sample": desc["sample"]}]
    th = reference_threads(cores)
    v, d, t = cpu_reference_rate(cfg, max(3.0, per_step), th)
    rows.append({"threads": th, "value": v, "unit": "MSamples/s", "seconds": t,
                 "row": "reference default floor(0.7*cores), main.cp:75", "sample": d["sample"]})
    ch = cfg["channels"]
    full_s = cfg["frames"] * ch / (value * 1e6)
    full = None
    if full_s <= 30.0:                                     # configs 1 and 2 timed in full
        v, d, t = cpu_reference_rate(cfg, 0.0, cores, frames=cfg["frames"])
        full = {"value": v, "unit": "MSamples/s", "seconds": t, "threads": cores, "sample": d["sample"]}
    else:
        full = {"value": None, "note": f"the whole file would take ~{full_s:.0f} s on {cores} threads: slice only"}
    world = int(os.environ.get("WORLD_SIZE", str(a.gpus)))
    line = {
        "impl": "reference", "metric": METRIC,
        "value": value, "unit": "MSamples/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": statistics.median(secs) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 samples x f64 taps (reference)", "data": "synthetic",
        "config": config_dict(cfg, max(world, a.gpus), a.mode),
        "cpu_baseline": {"value": value, "unit": "MSamples/s", "kind": desc["kind"], "cores": cores,
                         "sample": desc["sample"], "rows": rows, "whole_file": full,
                         "times": "PCM in -> PCM out: decode + FIR + peak + normalise + encode "
                                  f"(the FIR is {desc['fir_share'] * 100:.1f} % of it)"},
        "e2e": {"value": value, "unit": "MSamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


if __name__ == "__main__":
    sys.exit(main())
