"""CPU test of the N>1 path (world_size 2 and 3, gloo): the host-side logic of the
sample-block mode -- block plan with (taps-1) halo, all-reduce MAX of the peak, common
scale, per-rank encode -- and of the batch mode's file assignment.  The device is
replaced by an oracle-backed stand-in (test infrastructure; the product has no CPU
path), so what is verified is the sharding/collective logic: the assembled PCM must
equal the unsharded oracle result bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest

import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


class OracleDevice:
    """Duck-types capi.Context for audio_fir_filter_b200.dist.process_block_rank."""

    device = 0

    def __init__(self, oracle, fc, bw):
        self.o, self.fc, self.bw = oracle, fc, bw

    def apply(self, kernel, pcm, frames, channels, bits, be, halo_left=0, halo_right=0):
        self.fmt = (frames, channels, bits, be)
        if frames == 0:
            self.y, self.pk = np.zeros((channels, 0)), 0.0
            return
        r = self.o.process(pcm, frames, channels, bits, be, self.fc, self.bw, False, halo_left, halo_right)
        self.y, self.pk = r["y"], r["peak"]

    def peak(self):
        return self.pk

    def encode(self, scale, out):
        frames, channels, bits, be = self.fmt
        out[:] = self.o.encode(self.y, scale, bits, be)


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    from audio_fir_filter_b200.dist import assign_files, max_over_ranks, my_block, process_block_rank
    from audio_fir_filter_b200.process import FilterOptions, PcmInfo, block_view

    oracle.set_threads(1)
    fs, frames, ch, bits, be = 8000, 20_011, 2, 24, False
    opts = FilterOptions(freq=40.0, slope=100.0, normalize=True)
    info = PcmInfo(frames, ch, bits, be, float(fs))
    pcm = oracle.synth_pcm(77, 0, frames, ch, bits, be, fs)          # every rank can read "the file"
    H = oracle.kernel_order(opts.slope / fs) // 2
    blk = my_block(info, H)
    dev = OracleDevice(oracle, opts.freq / fs, opts.slope / fs)
    out = np.zeros(blk.frames * info.frame_bytes, dtype=np.uint8)
    r = process_block_rank(dev, None, block_view(pcm, info, blk), info, blk, opts, out)
    # every rank ends up with the same global peak and scale
    peaks = [None] * world
    dist.all_gather_object(peaks, (r["peak"], r["scale"], dev.pk))
    assert len({p[:2] for p in peaks}) == 1
    assert r["peak"] == max(p[2] for p in peaks)
    assert max_over_ranks(float(rank)) == world - 1
    parts = [None] * world
    dist.all_gather_object(parts, out)
    if rank == 0:
        whole = oracle.process(pcm, frames, ch, bits, be, opts.freq / fs, opts.slope / fs, True)
        got = np.concatenate(parts)
        assert whole["peak"] == r["peak"] and whole["scale"] == r["scale"]
        assert np.array_equal(got, whole["pcm"])
        # batch mode: every file lands on exactly one rank
        per = assign_files([10, 3, 8, 8, 1], world)
        assert sorted(i for lst in per for i in lst) == [0, 1, 2, 3, 4]
        open(os.path.join(tmp, "ok"), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sample_block_mode_over_gloo(tmp_path, oracle_mod, world):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert (tmp_path / "ok").exists()


def _bench_worker(rank, world, port, tmp):
    """bench.py's multi-rank plumbing over gloo: every rank checks the windows of ITS block with the
    parity checker (an oracle-backed stand-in serves the 'parked signal'), the rows are gathered,
    timings are max-reduced -- what `bench.py --gpus N` does around the GPU work."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    import oracle
    from test_bench_parity_checker import CFG, FakeCtx, FakeKernel, make_block

    oracle.set_threads(1)
    taps, blk, y = make_block(oracle, world, rank)
    if rank == world - 1 and os.path.exists(os.path.join(tmp, "inject")):
        y = y.copy()
        y[0, 1] += 1e-9                                   # a defect at the last rank's left seam
    pw = bench.ParityWindows(CFG, blk, CFG["frames"], bench.SEED, FakeKernel(taps), rank, W=256, n_random=2)
    pw.check_taps(CFG).check_signal(FakeCtx(y))
    r = pw.result
    rows = bench.gather_rows(world, "cpu", [r["windows"], r["worst_d3"], 1.0 if r["ok"] else 0.0, float(np.abs(y).max())])
    assert len(rows) == world and all(len(x) == 4 for x in rows)
    assert rows[rank][0] == r["windows"]
    (slowest,) = bench.max_over_ranks(world, "cpu", float(10 + rank))
    assert slowest == 10 + world - 1
    if rank == 0:
        verdict = all(x[2] == 1.0 for x in rows)
        open(os.path.join(tmp, "verdict"), "w").write("ok" if verdict else "fail")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("inject", [False, True])
def test_bench_parity_gather_over_gloo(tmp_path, oracle_mod, inject):
    if inject:
        (tmp_path / "inject").write_text("1")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_bench_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "verdict").read_text() == ("fail" if inject else "ok")
