"""Builders of small WAVE / RF64 / AIFF / AIFF-C files with awkward chunk layouts, for the
container and CLI tests: unknown chunks before and after the samples, odd-sized chunks
with their pad byte, WAVE_FORMAT_EXTENSIBLE, SSND with a non-zero offset, COMM after SSND."""
import struct


def _riff_chunk(cid: bytes, data: bytes) -> bytes:
    return cid + struct.pack("<I", len(data)) + data + (b"\0" if len(data) & 1 else b"")


def _iff_chunk(cid: bytes, data: bytes) -> bytes:
    return cid + struct.pack(">I", len(data)) + data + (b"\0" if len(data) & 1 else b"")


def wav_bytes(pcm: bytes, channels: int, bits: int, rate: int, extensible=False, extra_before=True, extra_after=True,
              rf64=False):
    nb = bits // 8
    if extensible:
        fmt = struct.pack("<HHIIHHHHIH", 0xFFFE, channels, rate, rate * channels * nb, channels * nb, bits, 22, bits,
                          (1 << channels) - 1, 1) + bytes.fromhex("000000001000800000aa00389b71")
    else:
        fmt = struct.pack("<HHIIHH", 1, channels, rate, rate * channels * nb, channels * nb, bits)
    body = b""
    if rf64:
        body += _riff_chunk(b"ds64", struct.pack("<QQQI", 0, len(pcm), len(pcm) // (channels * nb), 0))
    body += _riff_chunk(b"fmt ", fmt)
    if extra_before:
        body += _riff_chunk(b"bext", b"originator\0" * 7)           # odd length -> pad byte
        body += _riff_chunk(b"JUNK", b"\x01\x02\x03")
    if rf64:
        body += b"data" + struct.pack("<I", 0xFFFFFFFF) + pcm + (b"\0" if len(pcm) & 1 else b"")
    else:
        body += _riff_chunk(b"data", pcm)
    if extra_after:
        body += _riff_chunk(b"LIST", b"INFOINAM\x05\0\0\0test\0\0")
        body += _riff_chunk(b"id3 ", b"ID3\x03\0\0\0\0\0\x0bxyz" + b"\xff" * 5)
    if rf64:
        return b"RF64" + struct.pack("<I", 0xFFFFFFFF) + b"WAVE" + body
    return b"RIFF" + struct.pack("<I", 4 + len(body)) + b"WAVE" + body


def _ext80(rate: float) -> bytes:
    import math

    m, e = math.frexp(rate)          # rate = m * 2^e, 0.5 <= m < 1
    mant = int(m * (1 << 64))
    return struct.pack(">HQ", e - 1 + 16383, mant)


def aiff_bytes(pcm: bytes, channels: int, bits: int, rate: float, ssnd_offset=0, comm_last=False, aifc=None,
               extra=True):
    frames = len(pcm) // (channels * (bits // 8))
    comm = struct.pack(">hIh", channels, frames, bits) + _ext80(rate)
    if aifc:
        comm += aifc + b"\x0enot compressed\0"   # pstring (14 chars) + pad to even
    ssnd = struct.pack(">II", ssnd_offset, 0) + b"\xAB" * ssnd_offset + pcm
    parts = []
    if aifc:
        parts.append(_iff_chunk(b"FVER", struct.pack(">I", 0xA2805140)))
    if not comm_last:
        parts.append(_iff_chunk(b"COMM", comm))
    if extra:
        parts.append(_iff_chunk(b"NAME", b"odd"))                    # odd length
        parts.append(_iff_chunk(b"MARK", struct.pack(">H", 0)))
    parts.append(_iff_chunk(b"SSND", ssnd))
    if comm_last:
        parts.append(_iff_chunk(b"COMM", comm))
    if extra:
        parts.append(_iff_chunk(b"ANNO", b"after the samples"))
        parts.append(_iff_chunk(b"ID3 ", b"ID3\x04" + b"\0" * 9))
    body = b"".join(parts)
    return b"FORM" + struct.pack(">I", 4 + len(body)) + (b"AIFC" if aifc else b"AIFF") + body
