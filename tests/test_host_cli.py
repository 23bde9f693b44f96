"""CPU tests of the C++23 host: the chunk-preserving container layer (through
host/container_tool) and the CLI's validation / exit-code behaviour (main.cp:84-164).
No GPU is touched: every error below is raised before a device is needed."""
import json
import os
import subprocess
import wave

import numpy as np
import pytest

from audio_fixtures import aiff_bytes, wav_bytes
from conftest import ROOT

TOOL = os.path.join(ROOT, "host", "container_tool")
LOWCUT = os.path.join(ROOT, "host", "lowcut")


@pytest.fixture(scope="module", autouse=True)
def built():
    r = subprocess.run(["make", "-C", os.path.join(ROOT, "host")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def info(path):
    r = subprocess.run([TOOL, "info", str(path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return json.loads(r.stdout)


def rand_pcm(frames, channels, bits, seed=0):
    return np.random.default_rng(seed).integers(0, 256, frames * channels * bits // 8, dtype=np.uint8).tobytes()


CASES = [
    ("plain16.wav", lambda p: wav_bytes(p, 2, 16, 44100, extra_before=False, extra_after=False), 2, 16, False, 44100),
    ("chunks24.wav", lambda p: wav_bytes(p, 2, 24, 48000), 2, 24, False, 48000),
    ("ext32.wav", lambda p: wav_bytes(p, 6, 32, 96000, extensible=True), 6, 32, False, 96000),
    ("big.rf64.wav", lambda p: wav_bytes(p, 2, 24, 192000, rf64=True), 2, 24, False, 192000),
    ("plain16.aif", lambda p: aiff_bytes(p, 2, 16, 44100.0), 2, 16, True, 44100),
    ("offset24.aif", lambda p: aiff_bytes(p, 3, 24, 48000.0, ssnd_offset=6), 3, 24, True, 48000),
    ("commlast32.aif", lambda p: aiff_bytes(p, 1, 32, 88200.0, comm_last=True), 1, 32, True, 88200),
    ("sowt.aifc", lambda p: aiff_bytes(p, 2, 16, 22050.0, aifc=b"sowt"), 2, 16, False, 22050),
    ("none.aifc", lambda p: aiff_bytes(p, 2, 24, 48000.0, aifc=b"NONE"), 2, 24, True, 48000),
]


@pytest.mark.parametrize("name,build,ch,bits,be,rate", CASES)
def test_container_layout_and_byte_identity(tmp_path, name, build, ch, bits, be, rate):
    frames = 1001                                   # odd payload sizes for 24-bit mono/3ch: pad byte
    pcm = rand_pcm(frames, ch, bits, seed=len(name))
    data = build(pcm)
    src = tmp_path / name
    src.write_bytes(data)
    i = info(src)
    assert (i["channels"], i["bits"], i["big_endian"], i["rate"], i["frames"]) == (ch, bits, be, rate, frames)
    off, n = i["payload_offset"], i["payload_bytes"]
    assert data[off:off + n] == pcm
    assert i["file_size"] == len(data)
    assert len(i["chunks"]) >= 2
    # output = input outside the payload, byte for byte (ProcessFile.cp:107-117)
    dst = tmp_path / ("out_" + name)
    r = subprocess.run([TOOL, "invert", str(src), str(dst)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = dst.read_bytes()
    assert len(out) == len(data)
    assert out[:off] == data[:off] and out[off + n:] == data[off + n:]
    assert out[off:off + n] == bytes(b ^ 0xFF for b in pcm)


def test_container_agrees_with_python_wave_module(tmp_path):
    p = tmp_path / "std.wav"
    with wave.open(str(p), "wb") as w:
        w.setnchannels(2)
        w.setsampwidth(3)
        w.setframerate(48000)
        w.writeframes(rand_pcm(500, 2, 24))
    i = info(p)
    assert (i["type"], i["channels"], i["bits"], i["frames"], i["rate"]) == ("WAVE", 2, 24, 500, 48000)
    # and the other way: what the tool writes is still a file the stdlib reads
    q = tmp_path / "inv.wav"
    assert subprocess.run([TOOL, "invert", str(p), str(q)]).returncode == 0
    with wave.open(str(q), "rb") as w:
        assert (w.getnchannels(), w.getsampwidth(), w.getframerate(), w.getnframes()) == (2, 3, 48000, 500)


def test_container_truncated_and_rejected_formats(tmp_path):
    pcm = rand_pcm(100, 2, 16)
    data = wav_bytes(pcm, 2, 16, 44100, extra_before=False, extra_after=False)
    t = tmp_path / "trunc.wav"
    t.write_bytes(data[:-51])                        # data chunk claims more than the file holds
    assert info(t)["frames"] == (len(pcm) - 51) // 4
    f = tmp_path / "float.wav"
    f.write_bytes(data.replace(b"\x01\x00\x02\x00", b"\x03\x00\x02\x00", 1))   # format tag 3 = IEEE float
    r = subprocess.run([TOOL, "info", str(f)], capture_output=True, text=True)
    assert r.returncode == 1 and "not integer PCM" in r.stderr
    e = tmp_path / "eight.wav"
    e.write_bytes(wav_bytes(rand_pcm(10, 1, 8), 1, 8, 8000))
    r = subprocess.run([TOOL, "info", str(e)], capture_output=True, text=True)
    assert r.returncode == 1 and "16-, 24- and 32-bit" in r.stderr
    g = tmp_path / "garbage.wav"
    g.write_bytes(b"not an audio file at all")
    r = subprocess.run([TOOL, "info", str(g)], capture_output=True, text=True)
    assert r.returncode == 1 and "not a WAVE or AIFF" in r.stderr


# ---------------------------------------------------------------- CLI -------------

def run(*args):
    return subprocess.run([LOWCUT, *map(str, args)], capture_output=True, text=True)


def test_cli_help_goes_to_stdout_with_success():
    r = run("--help")                                 # StopNoError: main.cp:64-66,153-156
    assert r.returncode == 0 and "Usage:" in r.stdout and "--frequency" in r.stdout and r.stderr == ""
    assert run("-h").returncode == 0


def test_cli_parameter_count_and_unknown_options():
    r = run()
    assert r.returncode == 1 and "Invalid number of parameters. Need at least 2." in r.stderr
    assert run("only.wav").returncode == 1
    r = run("--bogus", "a.wav", "b.wav")
    assert r.returncode == 1 and "unrecognised option" in r.stderr
    r = run("-f")
    assert r.returncode == 1 and "missing" in r.stderr
    r = run("-f", "abc", "a.wav", "b.wav")
    assert r.returncode == 1 and "invalid" in r.stderr


def test_cli_single_file_scenario_validation(tmp_path):
    src = tmp_path / "in.wav"
    src.write_bytes(wav_bytes(rand_pcm(64, 2, 16), 2, 16, 44100))
    r = run(tmp_path / "missing.wav", tmp_path / "out.wav")
    assert r.returncode == 1 and "missing.wav" in r.stderr                       # FileNotFound
    d = tmp_path / "dir.wav"
    d.mkdir()
    r = run(src, d)
    assert r.returncode == 1 and "must be a file path, not a directory" in r.stderr
    r = run(src, tmp_path / "out.aif")
    assert r.returncode == 1 and "extensions must match" in r.stderr
    existing = tmp_path / "exists.wav"
    existing.write_bytes(b"keep me")
    r = run(src, existing)
    assert r.returncode == 1 and "exists.wav" in r.stderr                        # FileExists without -O
    assert existing.read_bytes() == b"keep me"
    r = run("-s", "0", src, tmp_path / "o.wav")
    assert r.returncode == 1 and "slope" in r.stderr


def test_cli_batch_scenario_validation(tmp_path):
    a = tmp_path / "a.wav"
    a.write_bytes(wav_bytes(rand_pcm(64, 2, 16), 2, 16, 44100))
    b = tmp_path / "b.wav"
    b.write_bytes(a.read_bytes())
    notdir = tmp_path / "file.txt"
    notdir.write_text("x")
    r = run(a, b, notdir)
    assert r.returncode == 1 and "Destination exists but is not a directory" in r.stderr
    r = run(a, b, tmp_path / "newdir.out")
    assert r.returncode == 1 and "does not exist and has a suffix" in r.stderr
    outdir = tmp_path / "out"
    outdir.mkdir()
    r = run(a, tmp_path / "nope.wav", outdir)
    assert r.returncode == 1 and "nope.wav" in r.stderr
    (outdir / "a.wav").write_bytes(b"old")
    r = run(a, b, outdir)
    assert r.returncode == 1 and "a.wav" in r.stderr                             # FileExists
    assert (outdir / "a.wav").read_bytes() == b"old"


def test_cli_without_a_gpu_fails_loudly_never_falls_back(tmp_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    src = tmp_path / "in.wav"
    src.write_bytes(wav_bytes(rand_pcm(64, 2, 16), 2, 16, 44100))
    r = run(src, tmp_path / "out.wav")
    assert r.returncode == 1 and "no CPU path" in r.stderr
    assert not (tmp_path / "out.wav").exists()


def test_container_parser_survives_malformed_files(tmp_path):
    """Truncations and corrupted size fields: the parser must answer (exit 0 or a clean error,
    exit 1) -- never crash, hang or read outside the file."""
    rng = np.random.default_rng(2026)
    seeds = [wav_bytes(rand_pcm(300, 2, 24), 2, 24, 48000), wav_bytes(rand_pcm(64, 2, 16), 2, 16, 44100, rf64=True),
             aiff_bytes(rand_pcm(300, 2, 16), 2, 16, 44100.0, ssnd_offset=2),
             aiff_bytes(rand_pcm(100, 1, 24), 1, 24, 48000.0, aifc=b"sowt", comm_last=True)]
    n = 0
    for si, data in enumerate(seeds):
        hdr_len = min(len(data), 200)
        variants = [data[:k] for k in (0, 3, 11, 12, 13, 20, 36, 43, 44, 45, len(data) // 2, len(data) - 1)]
        for _ in range(40):
            b = bytearray(data)
            for _ in range(int(rng.integers(1, 4))):
                pos = int(rng.integers(0, hdr_len))
                b[pos] = int(rng.integers(0, 256))
            variants.append(bytes(b))
        # size fields blown up / zeroed
        for pos in range(4, hdr_len - 4, 7):
            b = bytearray(data)
            b[pos:pos + 4] = b"\xff\xff\xff\xff" if pos % 2 else b"\0\0\0\0"
            variants.append(bytes(b))
        for vi, v in enumerate(variants):
            p = tmp_path / f"m{si}_{vi}.bin"
            p.write_bytes(v)
            r = subprocess.run([TOOL, "info", str(p)], capture_output=True, text=True, timeout=10)
            assert r.returncode in (0, 1), (si, vi, r.returncode, r.stderr[-200:])
            if r.returncode == 0:
                i = json.loads(r.stdout)
                assert i["payload_offset"] + i["payload_bytes"] <= len(v)
                assert i["frames"] * i["channels"] * i["bits"] // 8 == i["payload_bytes"]
                q = tmp_path / "o.bin"
                r2 = subprocess.run([TOOL, "invert", str(p), str(q)], capture_output=True, text=True, timeout=10)
                assert r2.returncode in (0, 1)
            n += 1
    assert n > 200


def _random_riff(rng, pcm, ch, bits, rate):
    """A WAVE file with random foreign chunks (random ids and sizes, odd ones padded) around fmt/data."""
    import struct

    nb = bits // 8

    def chunk(cid, data):
        return cid + struct.pack("<I", len(data)) + data + (b"\0" if len(data) & 1 else b"")

    def junk():
        cid = bytes(rng.integers(65, 91, 4).astype(np.uint8))          # 'A'..'Z'
        if cid in (b"DATA", b"FMT ", b"DSSF"):
            cid = b"XTRA"
        return chunk(cid, rng.integers(0, 256, int(rng.integers(0, 40)), dtype=np.uint8).tobytes())

    body = b"".join(junk() for _ in range(int(rng.integers(0, 4))))
    body += chunk(b"fmt ", struct.pack("<HHIIHH", 1, ch, rate, rate * ch * nb, ch * nb, bits))
    body += b"".join(junk() for _ in range(int(rng.integers(0, 4))))
    body += chunk(b"data", pcm)
    body += b"".join(junk() for _ in range(int(rng.integers(0, 4))))
    return b"RIFF" + struct.pack("<I", 4 + len(body)) + b"WAVE" + body


def test_container_random_chunk_layouts_keep_every_foreign_byte(tmp_path):
    rng = np.random.default_rng(7)
    for case in range(40):
        ch = int(rng.integers(1, 9))
        bits = int(rng.choice([16, 24, 32]))
        frames = int(rng.integers(0, 200))
        pcm = rng.integers(0, 256, frames * ch * bits // 8, dtype=np.uint8).tobytes()
        data = _random_riff(rng, pcm, ch, bits, 48000)
        src, dst = tmp_path / f"r{case}.wav", tmp_path / f"r{case}_o.wav"
        src.write_bytes(data)
        i = info(src)
        assert (i["channels"], i["bits"], i["frames"], i["payload_bytes"]) == (ch, bits, frames, len(pcm))
        off = i["payload_offset"]
        assert data[off:off + len(pcm)] == pcm
        r = subprocess.run([TOOL, "invert", str(src), str(dst)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        out = dst.read_bytes()
        assert out[:off] == data[:off] and out[off + len(pcm):] == data[off + len(pcm):]
        assert out[off:off + len(pcm)] == bytes(b ^ 0xFF for b in pcm)


def test_cli_refuses_to_overwrite_its_own_input(tmp_path):
    """The reference deletes an existing output before opening the input (main.cp:107), so
    `lowcut -O x.wav x.wav` would destroy x.wav; this host refuses (documented deviation)."""
    src = tmp_path / "x.wav"
    data = wav_bytes(rand_pcm(64, 2, 16), 2, 16, 44100)
    src.write_bytes(data)
    r = run("-O", src, src)
    assert r.returncode == 1 and "same file" in r.stderr
    assert src.read_bytes() == data
    other = tmp_path / "y.wav"
    other.write_bytes(data)
    r = run("-O", src, other, tmp_path)          # batch into the inputs' own directory
    assert r.returncode == 1 and "same file" in r.stderr
    assert src.read_bytes() == data and other.read_bytes() == data


def test_cli_option_syntax_variants(tmp_path):
    """boost::program_options' default style, which the reference's CLI gets through c_lib:
    bundled switches, attached short values, --name=value, unambiguous prefixes, `--`."""
    missing = tmp_path / "nope.wav"
    out = tmp_path / "o.wav"
    # every form parses: the run then fails on the missing input, not on the syntax
    for args in (["-nvO", "-f20", "-s", "10"], ["--freq=20", "--slope", "10", "--norm"], ["-f", "20", "-t", "3", "-g", "1"],
                 ["-Ons25", "--verb"], ["-f=20"]):
        r = run(*args, missing, out)
        assert r.returncode == 1 and r.stderr.startswith("File not found:") and "nope.wav" in r.stderr, (args, r.stderr)
    # after `--` everything is a path, even if it looks like an option
    r = run("-f", "20", "--", "-n", out)
    assert r.returncode == 1 and "File not found: -n" in r.stderr
    # errors
    assert "ambiguous" not in run("--n", missing, out).stderr          # --n -> --normalize (only match)
    r = run("--normalize=1", missing, out)
    assert r.returncode == 1 and "does not take any arguments" in r.stderr
    r = run("-t", "-1", missing, out)
    assert r.returncode == 1 and "invalid" in r.stderr
    r = run("-f", "1e", missing, out)
    assert r.returncode == 1 and "invalid" in r.stderr


def test_container_sizes_near_2_64_cannot_wrap_the_clamp(tmp_path):
    """Chunk sizes come from the file.  An RF64 ds64 data size or an SSND offset near 2^64 / 2^32
    must not wrap `offset + size` past the truncation clamp: the payload is what the
    file holds, never more."""
    import struct

    pcm = rand_pcm(200, 2, 16)
    good = wav_bytes(pcm, 2, 16, 44100, rf64=True, extra_before=False, extra_after=False)
    at = good.index(b"ds64") + 8 + 8                       # the 64-bit data size inside ds64
    for size in (2**64 - 1, 2**64 - 8, 2**64 - len(good), 2**63, 2**32 + 5):
        b = bytearray(good)
        b[at:at + 8] = struct.pack("<Q", size)
        p = tmp_path / f"ds64_{size}.wav"
        p.write_bytes(bytes(b))
        i = info(p)
        assert i["payload_offset"] + i["payload_bytes"] <= len(b)
        assert i["frames"] == 200 and i["payload_bytes"] == len(pcm)      # clamped to what is there
        r = subprocess.run([TOOL, "invert", str(p), str(tmp_path / "o.wav")], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    aif = aiff_bytes(pcm, 2, 16, 44100.0, extra=False)
    at = aif.index(b"SSND") + 8                            # SSND offset field
    for off in (2**32 - 1, 2**32 - 8, len(aif), 2**31):
        b = bytearray(aif)
        b[at:at + 4] = struct.pack(">I", off)
        p = tmp_path / f"ssnd_{off}.aif"
        p.write_bytes(bytes(b))
        r = subprocess.run([TOOL, "info", str(p)], capture_output=True, text=True)
        assert r.returncode == 1 and "SSND offset" in r.stderr, (off, r.stdout, r.stderr)


def test_container_constructor_does_not_leak_descriptors_on_format_errors(tmp_path):
    """A FormatError thrown by the constructor must close the file again (the destructor of a
    partially built object never runs).  lowcut opens every batch input up front, so a leak
    here would be one descriptor per rejected file."""
    bad = tmp_path / "bad.wav"
    bad.write_bytes(b"RIFF\x04\0\0\0WAVE")                 # no fmt, no data
    src = os.path.join(ROOT, "tests", "harness", "fd_leak_probe.cpp")
    exe = tmp_path / "fd_leak_probe"
    r = subprocess.run(["g++", "-std=c++23", "-O1", "-I", os.path.join(ROOT, "host"), "-o", str(exe), src,
                        os.path.join(ROOT, "host", "audio_container.cpp")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe), str(bad)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "leaked=0" in r.stdout, r.stdout


def test_cli_batch_two_inputs_with_one_basename(tmp_path):
    """a/x.wav b/x.wav out/ map to one destination.  The reference runs files one after the
    other: its second iteration finds the first one's output and throws FileExists
    (main.cp:140-142).  Files run concurrently here, so the clash is found up front -- before a
    GPU is needed -- instead of two lanes corrupting one .part file."""
    (tmp_path / "a").mkdir()
    (tmp_path / "b").mkdir()
    data = wav_bytes(rand_pcm(64, 2, 16), 2, 16, 44100)
    (tmp_path / "a" / "x.wav").write_bytes(data)
    (tmp_path / "b" / "x.wav").write_bytes(data)
    out = tmp_path / "out"
    out.mkdir()
    r = run(tmp_path / "a" / "x.wav", tmp_path / "b" / "x.wav", out)
    assert r.returncode == 1 and "File exists" in r.stderr and "x.wav" in r.stderr
    assert not list(out.iterdir())                         # nothing was started


def test_cli_overwrite_keeps_old_outputs_until_each_file_is_done(tmp_path):
    """-O no longer deletes every existing destination before the first file runs: the old file
    is replaced by the rename that completes its successor (the reference removes each output
    only when it reaches that file, main.cp:144).  Without a GPU the run fails -- and the old
    outputs must still be there."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the run would succeed")
    a = tmp_path / "a.wav"
    a.write_bytes(wav_bytes(rand_pcm(64, 2, 16), 2, 16, 44100))
    b = tmp_path / "b.wav"
    b.write_bytes(a.read_bytes())
    out = tmp_path / "out"
    out.mkdir()
    (out / "a.wav").write_bytes(b"old a")
    (out / "b.wav").write_bytes(b"old b")
    r = run("-O", a, b, out)
    assert r.returncode == 1
    assert (out / "a.wav").read_bytes() == b"old a" and (out / "b.wav").read_bytes() == b"old b"
    single = tmp_path / "single.wav"
    single.write_bytes(b"old single")
    r = run("-O", a, single)
    assert r.returncode == 1 and single.read_bytes() == b"old single"
