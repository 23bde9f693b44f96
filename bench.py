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

``value``  : filtered output MSamples/s of the whole job with the PCM resident in HBM
             (device timing, CUDA events on the launching stream, max over ranks).
``e2e``    : the same metric through the host-buffer C-ABI (fir_gpu_apply /
             fir_gpu_peak / fir_gpu_encode): pinned host PCM in, pinned host PCM out,
             both copies inside the timed region.
``roofline``: the FIR kernel's algorithmic FP64 FLOP over its CUDA-event time, against
             the DFMA-pipe peak measured live by a register-resident probe.
N > 1      : ONE long file of N x the payload, sample-block sharded with a (taps-1)
             halo; the only collective is the all-reduce MAX of the peak.
             ``--mode batch`` instead gives every rank its own files (no collective).

--impl reference times the CPU restatement of the reference's multithreaded path
(oracle/_ref = the reference's FilterCore.h compiled in place, else the oracle
port) on the host cores; that is the only use of oracle/ here besides cpu_baseline.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SEED = 0x00F1F1F1

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

def cpu_reference_rate(cfg: dict, budget_s: float, threads: int | None = None):
    """The reference's multithreaded CPU path (ProcessFile.cp:57-87 fan-out around
    FilterCore.h apply_filter_range, float32 buffers) on a bounded sample of the same
    workload.  Returns (MSamples/s, description dict)."""
    import oracle

    oracle.build()
    threads = threads or len(os.sched_getaffinity(0))
    fs = cfg["fs"]
    taps = oracle.build_lowcut(cfg["freq"] / fs, cfg["slope"] / fs)
    use_ref = oracle.ref_lib() is not None
    ch = cfg["channels"]

    def run(frames: int) -> float:
        pcm = oracle.synth_pcm(SEED, 0, frames, ch, cfg["bits"], cfg["be"], fs)
        x = oracle.decode(pcm, frames, ch, cfg["bits"], cfg["be"], dtype=np.float32)
        t0 = time.perf_counter()
        for c in range(ch):                                   # channels sequential (ProcessFile.cp:57)
            if use_ref:
                oracle.ref_filter_channel_threads(x[c], taps, threads)
            else:
                oracle.fir_f32(x[c], taps, threads=threads)
        return time.perf_counter() - t0

    # calibrate on a short file, then size the sample for the time budget
    n0 = max(4 * taps.size, 1 << 15)
    t = run(n0)
    rate = n0 * ch / t
    frames = int(min(cfg["frames"], max(n0, rate * budget_s / ch)))
    t = run(frames)
    msps = frames * ch / t / 1e6
    desc = {
        "kind": "reference" if use_ref else "port",
        "cores": threads,
        "sample": (f"first {frames} frames x {ch} ch of the workload ({frames * ch / 1e6:.2f} MSamples, {t:.1f} s), same "
                   f"{taps.size}-tap kernel; " +
                   ("reference FilterCore.h apply_filter_range compiled in place (c_lib's fms()/taps behind "
                    "interface shims), " if use_ref else "oracle port of FilterCore.h (ref_f32 mode), ") +
                   f"float32 buffers, {threads} std::threads per channel as ProcessFile.cp:64-83, -O3 " +
                   ("AVX-512" if use_ref and "avx512" in getattr(oracle.ref_lib(), "_path", "") else "AVX2/FMA")),
    }
    return msps, desc, t


# --------------------------------------------------------------------- main ----------

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
    ap.add_argument("--variant", type=int, default=-1, help="FIR kernel variant (experiments)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    a = ap.parse_args()
    if a.warmup < 3 and a.config != 6:
        a.warmup = 3 if a.impl == "ours" else a.warmup
    cfg = CONFIGS[a.config]
    inplace = a.config == 6          # the encoded PCM overwrites the input block: 180 GB would not hold both
    if inplace:
        a.steps, a.warmup, a.no_e2e, a.no_cpu = 1, 1, True, True
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if a.impl == "reference":
        return reference_arm(a, cfg, rank)

    import torch
    import torch.distributed as dist

    from audio_fir_filter_b200 import capi
    from audio_fir_filter_b200.dist import allreduce_max_peak
    from audio_fir_filter_b200.process import plan_blocks, scale_for_peak

    if not torch.cuda.is_available():
        print("bench.py: no CUDA device; this framework has no CPU path", file=sys.stderr)
        return 2
    numa_note = bind_to_gpu_numa_node(local) if world > 1 else "numa: not bound (single rank)"
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = torch.device(f"cuda:{local}")

    fs, ch, bits, be = cfg["fs"], cfg["channels"], cfg["bits"], cfg["be"]
    fb = ch * bits // 8
    taps = kernel_order(cfg["slope"] / fs) + 1
    H = (taps - 1) // 2
    per_rank = cfg["frames"]
    if a.mode == "block":
        total_frames = per_rank * world
        blk = plan_blocks(total_frames, world, H)[rank]
    else:
        total_frames = per_rank
        blk = plan_blocks(per_rank, 1, H)[0]
    first = blk.start - blk.halo_left
    n_in = blk.total_frames
    in_bytes, out_bytes = n_in * fb, blk.frames * fb

    ctx = capi.Context(local)
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)
    if a.variant >= 0:
        ctx.set_variant(a.variant)
    kernel = ctx.build_kernel(cfg["freq"] / fs, cfg["slope"] / fs)
    assert kernel.num_taps == taps

    # measured FP64 ceilings of this GPU, before the run (register-resident probes)
    dfma_peak = ctx.fp64_peak(0, 0.25)
    dmma_peak = ctx.fp64_peak(1, 0.25)

    seed = SEED if a.mode == "block" else SEED + rank
    if inplace and world != 8:
        print("bench.py: --config 6 (config 5 in full) needs --gpus 8", file=sys.stderr)
        return 2
    d_in = torch.empty(in_bytes, dtype=torch.uint8, device=dev)
    d_out = d_in[:out_bytes] if inplace else torch.empty(out_bytes, dtype=torch.uint8, device=dev)

    def synth():
        ctx.synth_pcm_dev(seed, first, n_in, ch, bits, be, fs, 1.0, d_in)
        ctx.synchronize()

    synth()

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    def reduce_peak() -> float:
        return allreduce_max_peak(ctx) if (world > 1 and a.mode == "block") else ctx.peak()

    def step_dev() -> float:
        ctx.apply_dev(kernel, d_in, blk.frames, ch, bits, be, blk.halo_left, blk.halo_right)
        pk = reduce_peak()
        ctx.encode_dev(scale_for_peak(pk, cfg["normalize"]), d_out)
        return pk

    # ---- kernel-resident arm (`value`) -------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None   # samples under load are picked by power draw
    for _ in range(a.warmup):
        step_dev()
        if inplace:
            synth()                  # the pass consumed its input
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fir_ms, dec_ms, enc_ms = [], [], []
    launches = 0
    e0.record(stream)
    t0 = time.perf_counter()
    for _ in range(a.steps):
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
    dev_ms = e0.elapsed_time(e1)
    tt = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms = float(tt[0]), float(tt[1])
    ms_per_step = dev_ms / a.steps
    out_samples_step = total_frames * ch if a.mode == "block" else per_rank * ch * world
    value = out_samples_step / (ms_per_step * 1e-3) / 1e6

    # ---- end-to-end arm through the host-buffer C-ABI ------------------------------------
    e2e = None
    if not a.no_e2e:
        h_in = capi.PinnedBuffer(in_bytes)
        h_out = capi.PinnedBuffer(out_bytes)
        h_in.array[:] = d_in.cpu().numpy()

        def step_host():
            if world == 1 or a.mode == "batch":
                # the one-call entry point a host makes per file (fir_gpu_process)
                p, _ = ctx.process(kernel, h_in.array, blk.frames, ch, bits, be, cfg["normalize"], h_out.array)
                return p
            ctx.apply(kernel, h_in.array, blk.frames, ch, bits, be, blk.halo_left, blk.halo_right)
            p = reduce_peak()
            ctx.encode(scale_for_peak(p, cfg["normalize"]), h_out.array)   # synchronous: D2H done on return
            return p

        for _ in range(a.warmup):
            step_host()
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            pk_host = step_host()
        barrier()
        e2e_ms = (time.perf_counter() - t0) * 1e3
        tt = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_ms = float(tt[0]) / a.steps
        th = ctx.last_timing()
        same = bool(np.array_equal(h_out.array, d_out.cpu().numpy())) and pk_host == pk
        e2e = {"value": out_samples_step / (e2e_ms * 1e-3) / 1e6, "unit": "MSamples/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes + 8,
               "h2d_ms": th["h2d_ms"], "d2h_ms": th["d2h_ms"], "host_buffers": "pinned (fir_gpu_host_alloc); " + numa_note,
               "matches_device_arm": same}
        h_in.free()
        h_out.free()

    # ---- roofline of the dominant kernel (FIR) -------------------------------------------
    flop_launch = algorithmic_flop(blk.frames, ch, taps, blk.halo_left, blk.halo_right)
    fir_avg_ms = sum(fir_ms) / len(fir_ms)
    achieved = flop_launch / (fir_avg_ms * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "fir_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(f"cfg{a.config}")
        except Exception:
            traffic = None
    vname = capi.variant_names()[a.variant if a.variant >= 0 else 0]
    is_dmma = vname.startswith("dmma")
    fp64_peak = dmma_peak if is_dmma else dfma_peak
    roofline = {
        "kernel": "fir_dmma_kernel" if is_dmma else "fir_fp64_kernel",
        # the FP64 tensor pipe (DMMA.8x8x4) for the default kernel, the FP64 FMA pipe for the dfma_* variants
        "bound": "tensor" if is_dmma else "fp64_fma",
        "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
        "frac": achieved / fp64_peak, "traffic": traffic,
        "peak_source": "measured live on this GPU before the run: register-resident "
                       + ("DMMA m8n8k4 f64 probe" if is_dmma else "DFMA probe")
                       + " (fir_gpu_fp64_peak); MEASURED_PEAKS.json has no FP64 figure (its bf16 number is for "
                         "tcgen05, which has no FP64 kind)",
        "dfma_probe": dfma_peak,
        "peak_nominal": 148 * 64 * 2 * 1.965e9 / 1e12, "dmma_probe": dmma_peak,
        "flop_per_launch": flop_launch, "fir_ms_per_launch": fir_avg_ms,
        "fir_share_of_step": fir_avg_ms / ms_per_step,
        "codec_hbm": {
            "decode_gbs": (n_in * fb + (blk.frames + 2 * H) * ch * 8) / (sum(dec_ms) / len(dec_ms) * 1e-3) / 1e9,
            "encode_gbs": (blk.frames * ch * 8 + out_bytes) / (sum(enc_ms) / len(enc_ms) * 1e-3) / 1e9,
            "peak_gbs": peak_hbm(),
        },
    }

    line = None
    if rank == 0:
        line = {
            "metric": "filtered output MSamples/s (FP64 direct FIR low-cut, PCM in -> PCM out)",
            "value": value, "unit": "MSamples/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": cfg["name"] + (f"; N={world}: one file {world}x as long, sample-block sharded with "
                                           f"{2 * H}-frame halo + NCCL all-reduce MAX of the peak"
                                           if world > 1 and a.mode == "block" else
                                           (f"; N={world}: one such file per rank (batch mode)" if world > 1 else "")),
                "taps": taps, "frames_per_gpu": blk.frames, "channels": ch, "bits": bits,
                "big_endian": be, "normalize": cfg["normalize"], "sample_rate": fs,
                "l2": "inputs larger than L2 (PCM + FP64 planes per step >> 126 MB)" if
                      (in_bytes + 16 * blk.frames * ch) > 200e6 else "working set may fit L2; see DESIGN.md",
                "fir_variant": vname,
            },
            "fp64_tflops": flop_launch * (world if a.mode == "block" else world) / (ms_per_step * 1e-3) / 1e12,
            "host_wall_ms_per_step": wall_ms / a.steps,
            "gpu_launches": launches,
            "clocks": clocks,
            "e2e": e2e,
            "roofline": roofline,
            "peak_value": pk,
        }
    kernel.free()

    if rank == 0 and not a.no_cpu and world == 1:   # the CPU leg is reported at N=1 only
        msps, desc, _ = cpu_reference_rate(cfg, a.cpu_seconds)
        line["cpu_baseline"] = {"value": msps, "unit": "MSamples/s", **desc}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(line)
    ctx.close()
    return 0


def peak_hbm() -> float:
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0  # B200_PROFILING.md fallback


def reference_arm(a, cfg, rank: int) -> int:
    """bench.py --impl reference: the reference's CPU implementation of the path on the
    host cores, rank 0 only; each step is a bounded sample of the workload."""
    if rank != 0:
        return 0
    import oracle

    oracle.build()
    threads = len(os.sched_getaffinity(0))
    # size one step for ~ (cpu-seconds / steps), at least 2 s
    per_step = max(2.0, a.cpu_seconds * 2.0 / max(1, a.steps))
    for _ in range(max(0, min(a.warmup, 1))):
        cpu_reference_rate(cfg, 1.0, threads)
    vals, secs, desc = [], [], None
    for _ in range(a.steps):
        v, desc, t = cpu_reference_rate(cfg, per_step, threads)
        vals.append(v)
        secs.append(t)
    value = statistics.median(vals)
    fs = cfg["fs"]
    taps = kernel_order(cfg["slope"] / fs) + 1
    line = {
        "impl": "reference",
        "metric": "filtered output MSamples/s (FP64 direct FIR low-cut, PCM in -> PCM out)",
        "value": value, "unit": "MSamples/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": statistics.median(secs) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 samples x f64 taps (reference)", "data": "synthetic",
        "config": {"workload": cfg["name"], "taps": taps, "channels": cfg["channels"], "bits": cfg["bits"],
                   "big_endian": cfg["be"], "sample_rate": fs},
        "cpu_baseline": {"value": value, "unit": "MSamples/s", **desc},
        "e2e": {"value": value, "unit": "MSamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


if __name__ == "__main__":
    sys.exit(main())
