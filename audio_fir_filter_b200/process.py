"""Host-side mirror of the arithmetic half of the reference's ``process_file``
(ProcessFile.cp:27-120) on top of the C-ABI -- the call sequence the C++ host
(``host/``) makes, restated in Python for the tests and ``bench.py``.

    FilterOptions              <- ProcessFile.h:13-19
    process_pcm                <- ProcessFile.cp:41-101,117  (decode .. encode of one payload)
    scale_for_peak             <- ProcessFile.cp:98-101      (auto-normalise rule)
    plan_blocks / process_pcm_sharded
                               <- SURVEY.md 8e: contiguous sample blocks with a
                                  (taps-1) halo, one max-reduction of the peak

Nothing here computes on the CPU: every sample goes through ``libfir_gpu.so``.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import capi


@dataclass
class FilterOptions:
    """``struct FilterOptions`` (ProcessFile.h:13-19).  ``num_threads`` is accepted for
    drop-in compatibility and ignored: the device grid replaces the thread fan-out."""

    freq: float = 15.0       # -f, Hz   (main.cp:44)
    slope: float = 10.0      # -s, Hz   (main.cp:46)
    normalize: bool = False  # -n       (main.cp:48)
    verbose: bool = False    # -v
    num_threads: int = 0     # -t


@dataclass
class PcmInfo:
    """What AudioFormat tells process_file about the sample chunk (ProcessFile.cp:35,43)."""

    frames: int
    channels: int
    bits: int
    big_endian: bool
    sample_rate: float

    @property
    def frame_bytes(self) -> int:
        return self.channels * self.bits // 8

    @property
    def nbytes(self) -> int:
        return self.frames * self.frame_bytes


def scale_for_peak(peak: float, normalize: bool) -> float:
    """ProcessFile.cp:98: ``if (maxMag > 1.0f || opts.normalize) normalize(buf)``;
    normalise = bring the peak to full scale (decision D2)."""
    if (peak > 1.0 or normalize) and peak > 0.0:
        return 1.0 / peak
    return 1.0


def process_pcm(ctx: capi.Context, pcm_in, info: PcmInfo, opts: FilterOptions, pcm_out,
                kernel: capi.Kernel | None = None) -> dict:
    """One payload through build_kernel -> apply -> peak -> scale rule -> encode.

    ``pcm_in`` / ``pcm_out`` are host byte buffers (numpy uint8 or addresses) of
    ``info.nbytes`` bytes.  Returns ``{"peak", "scale", "half_len", "taps"}``."""
    own = kernel is None
    if own:
        # ProcessFile.cp:48-49: both arguments are normalised by the sample rate.
        kernel = ctx.build_kernel(opts.freq / info.sample_rate, opts.slope / info.sample_rate)
    try:
        # one call: apply -> peak -> scale rule (ProcessFile.cp:98) -> encode, with the uploads
        # and (without -n) the downloads running under the FIR
        peak, scale = ctx.process(kernel, pcm_in, info.frames, info.channels, info.bits, info.big_endian,
                                  opts.normalize, pcm_out)
        assert scale == scale_for_peak(peak, opts.normalize)
        return {"peak": peak, "scale": scale, "half_len": kernel.half_len, "taps": kernel.num_taps}
    finally:
        if own:
            kernel.free()


# ---- sample-block sharding of one long payload ---------------------


@dataclass
class Block:
    """Frames ``[start, start+frames)`` of the file, owned by one GPU, plus the real
    halo frames the host hands over with it."""

    rank: int
    start: int
    frames: int
    halo_left: int
    halo_right: int

    @property
    def first_byte_frame(self) -> int:
        return self.start - self.halo_left

    @property
    def total_frames(self) -> int:
        return self.halo_left + self.frames + self.halo_right


def plan_blocks(total_frames: int, world: int, half_len: int, align: int = 16) -> list[Block]:
    """Contiguous blocks B = ceil(frames / world) rounded up to ``align`` frames; each block
    carries min(half_len, what exists) real frames either side -- implicit zeros
    only at the true file ends (FilterCore.h:57-61,72-76)."""
    if world < 1:
        raise ValueError("world must be >= 1")
    per = -(-total_frames // world)
    per = -(-per // align) * align if per else 0
    out = []
    for r in range(world):
        s = min(r * per, total_frames)
        e = min(s + per, total_frames)
        out.append(Block(r, s, e - s, min(half_len, s), min(half_len, total_frames - e)))
    return out


def process_block(ctx: capi.Context, kernel: capi.Kernel, pcm_block, info: PcmInfo, blk: Block) -> float:
    """Filter phase of one block: returns this device's peak (to be max-reduced)."""
    ctx.apply(kernel, pcm_block, blk.frames, info.channels, info.bits, info.big_endian,
              blk.halo_left, blk.halo_right)
    return ctx.peak() if blk.frames else 0.0


def block_view(pcm: np.ndarray, info: PcmInfo, blk: Block) -> np.ndarray:
    fb = info.frame_bytes
    return pcm[blk.first_byte_frame * fb:(blk.first_byte_frame + blk.total_frames) * fb]


def process_pcm_sharded(ctxs: list[capi.Context], pcm_in: np.ndarray, info: PcmInfo,
                        opts: FilterOptions) -> tuple[np.ndarray, dict]:
    """Single-process emulation of the sample-block mode over ``len(ctxs)`` contexts
    (one per GPU, or several on one GPU in tests): filter every block, max the
    peaks, encode every block with the common scale."""
    world = len(ctxs)
    kernels = [c.build_kernel(opts.freq / info.sample_rate, opts.slope / info.sample_rate) for c in ctxs]
    blocks = plan_blocks(info.frames, world, kernels[0].half_len)
    peaks = [process_block(ctxs[b.rank], kernels[b.rank], block_view(pcm_in, info, b), info, b)
             for b in blocks]
    peak = max(peaks)
    scale = scale_for_peak(peak, opts.normalize)
    out = np.empty(info.nbytes, dtype=np.uint8)
    fb = info.frame_bytes
    for b in blocks:
        if b.frames:
            ctxs[b.rank].encode(scale, out[b.start * fb:(b.start + b.frames) * fb])
    for k in kernels:
        k.free()
    return out, {"peak": peak, "scale": scale, "blocks": blocks, "peaks": peaks}
