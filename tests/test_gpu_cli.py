"""GPU tests of the C++23 host: `host/lowcut` end to end on real WAVE / AIFF files --
sample payload against the oracle, every other byte identical to the input."""
import os
import subprocess

import numpy as np
import pytest

from audio_fixtures import aiff_bytes, wav_bytes
from conftest import ROOT
from test_gpu_parity import lsb_flips

pytestmark = pytest.mark.gpu
LOWCUT = os.path.join(ROOT, "host", "lowcut")


@pytest.fixture(scope="module", autouse=True)
def built():
    if not os.path.exists(LOWCUT):
        r = subprocess.run(["make", "-C", os.path.join(ROOT, "host")], capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr


def run(*args):
    return subprocess.run([LOWCUT, *map(str, args)], capture_output=True, text=True)


def check_output(oracle_mod, src_bytes, out_bytes, pcm, ch, bits, be, fs, freq, slope, normalize):
    off = src_bytes.index(pcm[:64])
    n = len(pcm)
    assert len(out_bytes) == len(src_bytes)
    assert out_bytes[:off] == src_bytes[:off] and out_bytes[off + n:] == src_bytes[off + n:]   # non-audio bytes
    frames = n // (ch * bits // 8)
    want = oracle_mod.process(np.frombuffer(pcm, dtype=np.uint8), frames, ch, bits, be, freq / fs, slope / fs, normalize)
    got = np.frombuffer(out_bytes[off:off + n], dtype=np.uint8)
    nflip, mx = lsb_flips(got, want["pcm"], bits, be)
    assert mx <= 1 and nflip <= 2 + 1e-5 * got.size, (nflip, mx)
    return nflip


def test_single_file_wave_24bit_with_foreign_chunks(tmp_path, oracle_mod):
    fs, ch, bits, frames = 48000, 2, 24, 120_000
    pcm = oracle_mod.synth_pcm(0xF1F1F1, 0, frames, ch, bits, False, fs).tobytes()
    src = tmp_path / "in.wav"
    data = wav_bytes(pcm, ch, bits, fs)
    src.write_bytes(data)
    dst = tmp_path / "out.wav"
    r = run("-f", 20, "-s", 200, src, dst)
    assert r.returncode == 0, r.stderr
    assert "Processing file: in.wav" in r.stdout
    check_output(oracle_mod, data, dst.read_bytes(), pcm, ch, bits, False, fs, 20.0, 200.0, False)
    # existing output: refused without -O, replaced with it
    assert run("-f", 20, "-s", 200, src, dst).returncode == 1
    r = run("-O", "-v", "--frequency=20", "--slope", "200", "-n", src, dst)
    assert r.returncode == 0, r.stderr
    assert "Doing audio normalize." in r.stdout and "Writing output file." in r.stdout
    check_output(oracle_mod, data, dst.read_bytes(), pcm, ch, bits, False, fs, 20.0, 200.0, True)


def test_single_file_aiff_16bit_big_endian_normalize(tmp_path, oracle_mod):
    fs, ch, bits, frames = 44100, 2, 16, 100_000
    pcm = oracle_mod.synth_pcm(3, 0, frames, ch, bits, True, fs).tobytes()
    data = aiff_bytes(pcm, ch, bits, float(fs), ssnd_offset=4)
    src = tmp_path / "in.aif"
    src.write_bytes(data)
    dst = tmp_path / "out.aif"
    r = run("-f", 30, "-s", 300, "-n", src, dst)
    assert r.returncode == 0, r.stderr
    check_output(oracle_mod, data, dst.read_bytes(), pcm, ch, bits, True, fs, 30.0, 300.0, True)


def test_batch_mode_to_directory(tmp_path, oracle_mod):
    fs = 48000
    files = []
    for i, (ch, bits) in enumerate([(2, 24), (1, 16), (4, 32), (2, 24)]):
        pcm = oracle_mod.synth_pcm(100 + i, 0, 30_000 + 1000 * i, ch, bits, False, fs).tobytes()
        p = tmp_path / f"f{i}.wav"
        data = wav_bytes(pcm, ch, bits, fs, extensible=(bits == 32))
        p.write_bytes(data)
        files.append((p, data, pcm, ch, bits))
    outdir = tmp_path / "filtered"
    r = run("-f", 40, "-s", 400, *[f[0] for f in files], outdir)
    assert r.returncode == 0, r.stderr
    assert "Creating directory" in r.stdout
    for p, data, pcm, ch, bits in files:
        assert f"Processing file: {p.name}" in r.stdout
        check_output(oracle_mod, data, (outdir / p.name).read_bytes(), pcm, ch, bits, False, fs, 40.0, 400.0, False)


def test_sample_block_mode_across_gpus_equals_single_gpu(tmp_path, oracle_mod):
    from audio_fir_filter_b200 import capi

    if capi.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    fs, ch, bits, frames = 48000, 2, 24, 700_000
    pcm = oracle_mod.synth_pcm(5, 0, frames, ch, bits, False, fs).tobytes()
    src = tmp_path / "long.wav"
    src.write_bytes(wav_bytes(pcm, ch, bits, fs))
    one, two = tmp_path / "one.wav", tmp_path / "two.wav"
    assert run("-g", 1, "-f", 20, "-s", 100, "-n", src, one).returncode == 0
    # the peak exchange as the NCCL collective north_star names (forced: this file is far too short for
    # the communicator start-up to hide under its FIR, so the host would take the max itself)
    r = subprocess.run([LOWCUT, "-g", "2", "-v", "-f", "20", "-s", "100", "-n", str(src), str(two)], capture_output=True,
                       text=True, env=dict(os.environ, LOWCUT_NCCL="1"))
    assert r.returncode == 0 and "sample blocks" in r.stdout and "ncclAllReduce(max)" in r.stdout, r.stdout + r.stderr
    assert one.read_bytes() == two.read_bytes()
    # and the default for a short file (host max of the same scalars): same bytes
    three = tmp_path / "three.wav"
    r = run("-g", 2, "-v", "-f", 20, "-s", 100, "-n", src, three)
    assert r.returncode == 0 and "host max" in r.stdout
    assert one.read_bytes() == three.read_bytes()


def test_nccl_peak_allreduce_across_contexts_on_two_gpus(oracle_mod):
    """fir_gpu_allreduce_peak: two blocks of one file on two devices of this process; after the
    collective both contexts hold max(peak0, peak1) (ProcessFile.cp:92-96 over the whole file)."""
    from audio_fir_filter_b200 import Context, capi, plan_blocks

    if capi.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    fs, ch, bits, frames = 8000, 2, 16, 100_000
    pcm = oracle_mod.synth_pcm(11, 0, frames, ch, bits, True, fs)
    ctxs = [Context(0), Context(1)]
    capi.comm_prepare(ctxs)
    peaks = []
    fb = ch * bits // 8
    for r, cx in enumerate(ctxs):
        k = cx.build_kernel(40.0 / fs, 50.0 / fs)
        b = plan_blocks(frames, 2, k.half_len)[r]
        cx.apply(k, pcm[(b.start - b.halo_left) * fb:(b.start + b.frames + b.halo_right) * fb], b.frames, ch, bits, True,
                 b.halo_left, b.halo_right)
        peaks.append(cx.peak())
    g = capi.allreduce_peak(ctxs)
    assert g == max(peaks) and peaks[0] != peaks[1]
    assert ctxs[0].peak() == g and ctxs[1].peak() == g
    want = oracle_mod.process(pcm, frames, ch, bits, True, 40.0 / fs, 50.0 / fs, True)
    assert abs(g - want["peak"]) <= 1e-12 * want["peak"]
    for cx in ctxs:
        cx.close()


def test_batch_overwrite_with_one_basename_twice_last_input_wins(tmp_path, oracle_mod):
    """-O with a/x.wav b/x.wav out/: the reference processes both in turn, the second replacing
    the first's output (main.cp:143-146); here the clash is settled up front -- one job, the
    later input -- instead of two lanes writing one file."""
    fs, ch, bits, frames = 8000, 2, 16, 30_000
    (tmp_path / "a").mkdir()
    (tmp_path / "b").mkdir()
    pa = oracle_mod.synth_pcm(1, 0, frames, ch, bits, False, fs).tobytes()
    pb = oracle_mod.synth_pcm(2, 0, frames, ch, bits, False, fs).tobytes()
    (tmp_path / "a" / "x.wav").write_bytes(wav_bytes(pa, ch, bits, fs))
    db = wav_bytes(pb, ch, bits, fs)
    (tmp_path / "b" / "x.wav").write_bytes(db)
    out = tmp_path / "out"
    r = run("-O", "-f", 40, "-s", 400, tmp_path / "a" / "x.wav", tmp_path / "b" / "x.wav", out)
    assert r.returncode == 0, r.stderr
    assert sorted(p.name for p in out.iterdir()) == ["x.wav"]
    check_output(oracle_mod, db, (out / "x.wav").read_bytes(), pb, ch, bits, False, fs, 40.0, 400.0, False)


def test_config1_full_size_file_through_the_cli(tmp_path, oracle_mod):
    """BASELINE config 1 as a real file: 60 s stereo 48 kHz 24-bit WAVE with foreign chunks
    either side of `data`, `lowcut -f 20 -s 20` (9601 taps): the whole payload against
    oracle_process, every other byte against the input."""
    fs, ch, bits, frames = 48000, 2, 24, 2_880_000
    pcm = oracle_mod.synth_pcm(0xF1F1F1, 0, frames, ch, bits, False, fs).tobytes()
    data = wav_bytes(pcm, ch, bits, fs)
    src, dst = tmp_path / "cfg1.wav", tmp_path / "cfg1_out.wav"
    src.write_bytes(data)
    r = run("-f", 20, "-s", 20, src, dst)
    assert r.returncode == 0, r.stderr
    nflip = check_output(oracle_mod, data, dst.read_bytes(), pcm, ch, bits, False, fs, 20.0, 20.0, False)
    print(f"config 1 through the CLI: {nflip} of {frames * ch} samples differ from the oracle by 1 LSB")


@pytest.mark.parametrize("kind", ["rf64", "aifc_sowt", "aifc_none32", "aiff_comm_last"])
def test_other_containers_end_to_end(tmp_path, oracle_mod, kind):
    """RF64 (ds64 size fields), AIFF-C 'sowt' (little-endian) and 'NONE', COMM after SSND:
    the layout the container layer reports is the one the device decodes with."""
    frames = 50_000
    if kind == "rf64":
        fs, ch, bits, be, ext = 96000, 2, 24, False, ".wav"
        build = lambda p: wav_bytes(p, ch, bits, fs, rf64=True)                      # noqa: E731
    elif kind == "aifc_sowt":
        fs, ch, bits, be, ext = 44100, 2, 16, False, ".aifc"
        build = lambda p: aiff_bytes(p, ch, bits, float(fs), aifc=b"sowt")            # noqa: E731
    elif kind == "aifc_none32":
        fs, ch, bits, be, ext = 48000, 3, 32, True, ".aifc"
        build = lambda p: aiff_bytes(p, ch, bits, float(fs), aifc=b"NONE", ssnd_offset=8)   # noqa: E731
    else:
        fs, ch, bits, be, ext = 48000, 1, 24, True, ".aif"
        build = lambda p: aiff_bytes(p, ch, bits, float(fs), comm_last=True)          # noqa: E731
    pcm = oracle_mod.synth_pcm(hash(kind) & 0xFFFF, 0, frames, ch, bits, be, fs).tobytes()
    data = build(pcm)
    src, dst = tmp_path / ("in" + ext), tmp_path / ("out" + ext)
    src.write_bytes(data)
    r = run("-f", 25, "-s", 250, "-n", src, dst)
    assert r.returncode == 0, r.stderr
    check_output(oracle_mod, data, dst.read_bytes(), pcm, ch, bits, be, fs, 25.0, 250.0, True)


def test_batch_mode_one_process_per_gpu_equals_one_gpu(tmp_path, oracle_mod):
    """Batch scenario on several GPUs = one worker process per GPU pulling files from a shared
    counter (main.cp:132-147 is a plain loop over independent files: no communication).  Same
    bytes as on one GPU; a file that fails makes the run exit 1 with the reason on stderr once,
    and leaves no .part file behind."""
    from audio_fir_filter_b200 import capi

    if capi.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    fs, ch, bits = 48000, 2, 24
    files = []
    for i in range(7):
        frames = 40_000 + 5_000 * i
        pcm = oracle_mod.synth_pcm(100 + i, 0, frames, ch, bits, False, fs).tobytes()
        p = tmp_path / f"in{i}.wav"
        p.write_bytes(wav_bytes(pcm, ch, bits, fs))
        files.append(p)
    one, many = tmp_path / "one", tmp_path / "many"
    r1 = run("-g", 1, "-f", 40, "-s", 400, *files, one)
    assert r1.returncode == 0, r1.stderr
    r2 = run("-g", 2, "-f", 40, "-s", 400, *files, many)
    assert r2.returncode == 0, r2.stderr
    assert "Using up to 2 GPU(s)." in r2.stdout
    assert sorted(l for l in r2.stdout.splitlines() if l.startswith("Processing file:")) == \
        sorted(f"Processing file: {p.name}" for p in files)
    for p in files:
        assert (one / p.name).read_bytes() == (many / p.name).read_bytes()
    # one broken input: exit 1, its reason on stderr exactly once, nothing half-written left behind
    bad = tmp_path / "broken.wav"
    bad.write_bytes(b"RIFF\x04\0\0\0WAVE")
    out3 = tmp_path / "third"
    r3 = run("-g", 2, "-f", 40, "-s", 400, files[0], bad, files[1], files[2], out3)
    assert r3.returncode == 1
    assert r3.stderr.count("broken.wav") == 1, r3.stderr
    assert not [p for p in out3.iterdir() if p.name.endswith(".part")]
    # a failure INSIDE a worker process (the destination of one file is a directory: the rename that
    # completes it fails): exit 1, reason on stderr, no .part left, the other workers stop handing out
    out4 = tmp_path / "fourth"
    (out4 / files[3].name).mkdir(parents=True)
    r4 = run("-O", "-g", 2, "-f", 40, "-s", 400, *files, out4)
    assert r4.returncode == 1 and files[3].name in r4.stderr, r4.stderr
    assert not [p for p in out4.iterdir() if p.name.endswith(".part")]


def test_batch_worker_processes_on_one_gpu(tmp_path, oracle_mod):
    """The process-per-GPU batch path on a ONE-GPU box: LOWCUT_WORKER_DEVICES=0,0 (test hook) starts
    two worker processes that share GPU 0 -- fork before CUDA, job counter and failure flag in
    shared memory, each worker its own context.  Same bytes as the single-process run."""
    fs, ch, bits = 48000, 2, 24
    files = []
    for i in range(9):
        frames = 30_000 + 7_000 * i
        pcm = oracle_mod.synth_pcm(200 + i, 0, frames, ch, bits, False, fs).tobytes()
        p = tmp_path / f"w{i}.wav"
        p.write_bytes(wav_bytes(pcm, ch, bits, fs))
        files.append((p, pcm))
    one, two = tmp_path / "one", tmp_path / "two"
    r1 = subprocess.run([LOWCUT, "-f", "40", "-s", "400", *[str(p) for p, _ in files], str(one)], capture_output=True, text=True,
                        env=dict(os.environ, LOWCUT_SINGLE_PROCESS="1"))
    assert r1.returncode == 0, r1.stderr
    r2 = subprocess.run([LOWCUT, "-g", "2", "-f", "40", "-s", "400", *[str(p) for p, _ in files], str(two)], capture_output=True,
                        text=True, env=dict(os.environ, LOWCUT_WORKER_DEVICES="0,0"))
    assert r2.returncode == 0, r2.stderr
    assert "Using up to 2 GPU(s)." in r2.stdout
    assert sorted(l for l in r2.stdout.splitlines() if l.startswith("Processing file:")) == \
        sorted(f"Processing file: {p.name}" for p, _ in files)
    for p, pcm in files:
        assert (one / p.name).read_bytes() == (two / p.name).read_bytes()
    p, pcm = files[4]
    check_output(oracle_mod, p.read_bytes(), (two / p.name).read_bytes(), pcm, ch, bits, False, fs, 40.0, 400.0, False)
    # a worker whose device cannot be used (ordinal 99 does not exist) steps aside; the other one does every file
    four = tmp_path / "four"
    r4 = subprocess.run([LOWCUT, "-g", "2", "-f", "40", "-s", "400", *[str(p) for p, _ in files], str(four)], capture_output=True,
                        text=True, env=dict(os.environ, LOWCUT_WORKER_DEVICES="0,99"))
    assert r4.returncode == 0, r4.stderr
    for p, _ in files:
        assert (one / p.name).read_bytes() == (four / p.name).read_bytes()
    # no worker with a usable device at all: the run fails loudly (no CPU path)
    r5 = subprocess.run([LOWCUT, "-g", "2", "-f", "40", "-s", "400", *[str(p) for p, _ in files], str(tmp_path / "five")],
                        capture_output=True, text=True, env=dict(os.environ, LOWCUT_WORKER_DEVICES="98,99"))
    assert r5.returncode == 1 and "no CPU path" in r5.stderr
    # a failure inside a worker: exit 1, the reason on stderr, no .part left
    out3 = tmp_path / "three"
    (out3 / files[2][0].name).mkdir(parents=True)
    r3 = subprocess.run([LOWCUT, "-O", "-g", "2", "-f", "40", "-s", "400", *[str(p) for p, _ in files], str(out3)],
                        capture_output=True, text=True, env=dict(os.environ, LOWCUT_WORKER_DEVICES="0,0"))
    assert r3.returncode == 1 and files[2][0].name in r3.stderr, r3.stderr
    assert not [q for q in out3.iterdir() if q.name.endswith(".part")]
