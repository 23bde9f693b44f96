"""One process per GPU: the multi-GPU modes of SURVEY.md section 8e on torch.distributed.

  * batch mode (config 4)        -- files are independent (main.cp:132-147 is a plain
                                    loop): ``assign_files`` deals whole files to
                                    ranks, no data-path collective at all;
  * sample-block mode (config 5) -- rank r owns frames [r*B, (r+1)*B) of every
                                    channel plus a (taps-1)/2 halo either side
                                    (``process.plan_blocks``); the ONLY exchange is
                                    the max-reduction of the 8-byte peak
                                    (ProcessFile.cp:92-96 is a max over the whole
                                    file), after which every rank encodes its block
                                    with the same scale.

torch.distributed is plumbing: NCCL over NVLink on the GPU box (the all-reduce
runs on the device scalar the FIR epilogue wrote, ordered on the context's
stream), gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .process import Block, FilterOptions, PcmInfo, plan_blocks, scale_for_peak


class _DevScalar:
    """Zero-copy view of one FP64 in device memory for torch (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int):
        self.__cuda_array_interface__ = {"shape": (1,), "typestr": "<f8", "data": (ptr, False), "version": 3,
                                         "strides": None}


def allreduce_max_peak(ctx, peak_local: float | None = None, group=None) -> float:
    """Max over ranks of the per-device peak.  With NCCL the reduction runs in place
    on the device scalar (``fir_gpu_peak_dev``); with gloo on a host double."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return ctx.peak() if peak_local is None else peak_local
    if dist.get_backend(group) == "nccl":
        # the scalar is written on the context's stream, NCCL runs on torch's: join them
        ctx.synchronize()
        t = torch.as_tensor(_DevScalar(ctx.peak_dev()), device=f"cuda:{ctx.device}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        return float(t.item())
    v = ctx.peak() if peak_local is None else peak_local
    t = torch.tensor([v], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def my_block(info: PcmInfo, half_len: int, rank: int | None = None, world: int | None = None) -> Block:
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    return plan_blocks(info.frames, world, half_len)[rank]


def process_block_rank(ctx, kernel, pcm_block, info: PcmInfo, blk: Block, opts: FilterOptions, out_block,
                       group=None) -> dict:
    """This rank's share of one long file: filter the block, all-reduce the peak,
    encode with the common scale.  ``pcm_block`` starts at frame ``blk.start - blk.halo_left``;
    ``out_block`` receives ``blk.frames`` frames."""
    ctx.apply(kernel, pcm_block, blk.frames, info.channels, info.bits, info.big_endian, blk.halo_left,
              blk.halo_right)
    peak = allreduce_max_peak(ctx, None, group)
    scale = scale_for_peak(peak, opts.normalize)
    if blk.frames:
        ctx.encode(scale, out_block)
    return {"peak": peak, "scale": scale}


def assign_files(sizes: list[int], world: int) -> list[list[int]]:
    """Batch mode: whole files to ranks, greedy by size (largest first onto the least
    loaded rank).  Returns the file indices per rank; deterministic on every rank."""
    order = sorted(range(len(sizes)), key=lambda i: (-sizes[i], i))
    load = [0] * world
    out: list[list[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += sizes[i]
    for lst in out:
        lst.sort()
    return out


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Timing helper: max over ranks of a host float."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def gather_bytes(local: np.ndarray, group=None) -> list[np.ndarray] | None:
    """Test helper: collect every rank's output block on rank 0 (gloo)."""
    world = dist.get_world_size(group)
    objs = [None] * world if dist.get_rank(group) == 0 else None
    dist.gather_object(local, objs, dst=0, group=group)
    return objs
