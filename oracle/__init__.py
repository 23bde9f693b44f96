"""CPU oracle for the lowcut hot path -- TEST INFRASTRUCTURE ONLY.

ctypes front end of ``oracle/fir_oracle.c`` (a long-double restatement of the
reference's FilterCore.h / ProcessFile.cp arithmetic; see that file's header
for the citations and for why parity is "unpinned") and, where it was built, of
``oracle/_ref/libref_filtercore.so`` (the reference's own FilterCore.h compiled
against interface shims).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product
(``audio_fir_filter_b200``, ``host/``) must never do so.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_REF_PATH = os.path.join(_HERE, "_ref", "libref_filtercore.so")
_REF_PATH_AVX512 = os.path.join(_HERE, "_ref", "libref_filtercore_avx512.so")

_i64 = C.c_int64
_dp = C.POINTER(C.c_double)
_fp = C.POINTER(C.c_float)
_bp = C.POINTER(C.c_uint8)


def build(force: bool = False) -> None:
    """Compile liboracle.so (and _ref/ when /root/reference is present)."""
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "fir_oracle.c"))
    ):
        subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, capture_output=True)
    if os.path.exists("/root/reference/FilterCore.h") and (
            force or not os.path.exists(_REF_PATH) or not os.path.exists(_REF_PATH_AVX512)):
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, capture_output=True)


_lib = None
_ref = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        L.oracle_kernel_order.restype = _i64
        L.oracle_kernel_order.argtypes = [C.c_double]
        L.oracle_build_lowcut.restype = C.c_int
        L.oracle_build_lowcut.argtypes = [C.c_double, C.c_double, _dp, _i64, C.c_void_p]
        for name in ("oracle_fir_hi", "oracle_fir_abs_scale"):
            f = getattr(L, name)
            f.restype = None
            f.argtypes = [_dp, _i64, _dp, _i64, _dp, _i64, _i64]
        L.oracle_fir_f32_range.restype = None
        L.oracle_fir_f32_range.argtypes = [_fp, _i64, _dp, _i64, _fp, _i64, _i64]
        L.oracle_fir_f32_threads.restype = C.c_int
        L.oracle_fir_f32_threads.argtypes = [_fp, _i64, _dp, _i64, _fp, _i64, _i64, C.c_int]
        L.oracle_decode_f64.restype = C.c_int
        L.oracle_decode_f64.argtypes = [_bp, _i64, C.c_int, C.c_int, C.c_int, _dp, _i64]
        L.oracle_decode_f32.restype = C.c_int
        L.oracle_decode_f32.argtypes = [_bp, _i64, C.c_int, C.c_int, C.c_int, _fp, _i64]
        L.oracle_peak.restype = C.c_double
        L.oracle_peak.argtypes = [_dp, _i64, C.c_int, _i64]
        L.oracle_scale_for_peak.restype = C.c_double
        L.oracle_scale_for_peak.argtypes = [C.c_double, C.c_int]
        L.oracle_encode.restype = C.c_int
        L.oracle_encode.argtypes = [_dp, _i64, C.c_int, _i64, C.c_double, C.c_int, C.c_int, _bp]
        L.oracle_encode_f32.restype = C.c_int
        L.oracle_encode_f32.argtypes = [_fp, _i64, C.c_int, _i64, C.c_float, C.c_int, C.c_int, _bp]
        L.oracle_synth_sample.restype = _i64
        L.oracle_synth_sample.argtypes = [C.c_uint64, C.c_int, _i64, C.c_int, _i64, C.c_double]
        L.oracle_synth_pcm.restype = C.c_int
        L.oracle_synth_pcm.argtypes = [C.c_uint64, _i64, _i64, C.c_int, C.c_int, C.c_int, _i64,
                                       C.c_double, _bp]
        L.oracle_process.restype = C.c_int
        L.oracle_process.argtypes = [_bp, _i64, C.c_int, C.c_int, C.c_int, _i64, _i64, C.c_double,
                                     C.c_double, C.c_int, C.c_double, _bp, _dp, _dp, _dp]
        L.oracle_set_threads.restype = None
        L.oracle_set_threads.argtypes = [C.c_int]
        L.oracle_hw_threads.restype = C.c_int
        _lib = L
    return _lib


def ref_lib():
    """The reference's FilterCore.h build, or None where it was never built."""
    global _ref
    if _ref is None and os.path.exists(_REF_PATH):
        path = _REF_PATH
        try:  # the AVX-512 build of the same source where the host CPU has it (the CPU baseline's best shot)
            if os.path.exists(_REF_PATH_AVX512) and " avx512f " in open("/proc/cpuinfo").read():
                path = _REF_PATH_AVX512
        except OSError:
            pass
        R = C.CDLL(path)
        R._path = path
        R.ref_apply_filter_range.restype = None
        R.ref_apply_filter_range.argtypes = [_fp, C.c_longlong, _dp, C.c_longlong, _fp,
                                             C.c_longlong, C.c_longlong]
        R.ref_filter_channel_threads.restype = None
        R.ref_filter_channel_threads.argtypes = [_fp, C.c_longlong, _dp, C.c_longlong, _fp,
                                                 C.c_uint]
        _ref = R
    return _ref


def _d(a):
    return a.ctypes.data_as(_dp)


def _f(a):
    return a.ctypes.data_as(_fp)


def _b(a):
    return a.ctypes.data_as(_bp)


def hw_threads() -> int:
    return lib().oracle_hw_threads()


def set_threads(n: int) -> None:
    lib().oracle_set_threads(int(n))


def kernel_order(bw_norm: float) -> int:
    return int(lib().oracle_kernel_order(bw_norm))


def build_lowcut(fc_norm: float, bw_norm: float, want_ld: bool = False):
    """Taps h[0..M] as float64 (and optionally the raw 80-bit values as
    np.longdouble)."""
    M = kernel_order(bw_norm)
    if M < 0:
        raise ValueError("bad transition width")
    taps = np.empty(M + 1, dtype=np.float64)
    ld = np.empty(M + 1, dtype=np.longdouble) if want_ld else None
    rc = lib().oracle_build_lowcut(fc_norm, bw_norm, _d(taps), M + 1,
                                   ld.ctypes.data if want_ld else None)
    if rc:
        raise ValueError(f"oracle_build_lowcut rc={rc}")
    return (taps, ld) if want_ld else taps


def fir_hi(x: np.ndarray, taps: np.ndarray, start: int = 0, end: int | None = None) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float64)
    taps = np.ascontiguousarray(taps, dtype=np.float64)
    end = x.size if end is None else end
    y = np.zeros(x.size, dtype=np.float64)
    lib().oracle_fir_hi(_d(x), x.size, _d(taps), taps.size - 1, _d(y), start, end)
    return y


def fir_abs_scale(x: np.ndarray, taps: np.ndarray, start: int = 0, end: int | None = None):
    x = np.ascontiguousarray(x, dtype=np.float64)
    taps = np.ascontiguousarray(taps, dtype=np.float64)
    end = x.size if end is None else end
    s = np.zeros(x.size, dtype=np.float64)
    lib().oracle_fir_abs_scale(_d(x), x.size, _d(taps), taps.size - 1, _d(s), start, end)
    return s


def fir_f32(x: np.ndarray, taps: np.ndarray, start: int = 0, end: int | None = None,
            threads: int = 0) -> np.ndarray:
    x = np.ascontiguousarray(x, dtype=np.float32)
    taps = np.ascontiguousarray(taps, dtype=np.float64)
    end = x.size if end is None else end
    y = np.zeros(x.size, dtype=np.float32)
    if threads > 0:
        lib().oracle_fir_f32_threads(_f(x), x.size, _d(taps), taps.size - 1, _f(y), start, end,
                                     threads)
    else:
        lib().oracle_fir_f32_range(_f(x), x.size, _d(taps), taps.size - 1, _f(y), start, end)
    return y


def ref_fir_f32(x: np.ndarray, taps: np.ndarray, start: int = 0, end: int | None = None):
    R = ref_lib()
    if R is None:
        raise RuntimeError("oracle/_ref was not built")
    x = np.ascontiguousarray(x, dtype=np.float32)
    taps = np.ascontiguousarray(taps, dtype=np.float64)
    end = x.size if end is None else end
    y = np.zeros(x.size, dtype=np.float32)
    R.ref_apply_filter_range(_f(x), x.size, _d(taps), taps.size, _f(y), start, end)
    return y


def ref_filter_channel_threads(x: np.ndarray, taps: np.ndarray, threads: int) -> np.ndarray:
    R = ref_lib()
    if R is None:
        raise RuntimeError("oracle/_ref was not built")
    x = np.ascontiguousarray(x, dtype=np.float32)
    taps = np.ascontiguousarray(taps, dtype=np.float64)
    y = np.zeros(x.size, dtype=np.float32)
    R.ref_filter_channel_threads(_f(x), x.size, _d(taps), taps.size, _f(y), threads)
    return y


def decode(pcm: np.ndarray, frames: int, channels: int, bits: int, big_endian: bool,
           dtype=np.float64) -> np.ndarray:
    pcm = np.ascontiguousarray(pcm, dtype=np.uint8)
    assert pcm.size == frames * channels * bits // 8
    out = np.empty((channels, frames), dtype=dtype)
    if dtype == np.float64:
        rc = lib().oracle_decode_f64(_b(pcm), frames, channels, bits, int(big_endian), _d(out), frames)
    else:
        rc = lib().oracle_decode_f32(_b(pcm), frames, channels, bits, int(big_endian), _f(out), frames)
    if rc:
        raise ValueError("unsupported bit depth")
    return out


def peak(y: np.ndarray) -> float:
    y = np.ascontiguousarray(y, dtype=np.float64)
    ch, fr = y.shape
    return float(lib().oracle_peak(_d(y), fr, ch, fr))


def scale_for_peak(pk: float, normalize: bool) -> float:
    return float(lib().oracle_scale_for_peak(pk, int(normalize)))


def encode(y: np.ndarray, scale: float, bits: int, big_endian: bool) -> np.ndarray:
    y = np.ascontiguousarray(y, dtype=np.float64)
    ch, fr = y.shape
    pcm = np.empty(fr * ch * bits // 8, dtype=np.uint8)
    rc = lib().oracle_encode(_d(y), fr, ch, fr, scale, bits, int(big_endian), _b(pcm))
    if rc:
        raise ValueError("unsupported bit depth")
    return pcm


def encode_f32(y: np.ndarray, scale: float, bits: int, big_endian: bool) -> np.ndarray:
    y = np.ascontiguousarray(y, dtype=np.float32)
    ch, fr = y.shape
    pcm = np.empty(fr * ch * bits // 8, dtype=np.uint8)
    lib().oracle_encode_f32(_f(y), fr, ch, fr, scale, bits, int(big_endian), _b(pcm))
    return pcm


def synth_pcm(seed: int, first_frame: int, frames: int, channels: int, bits: int,
              big_endian: bool, rate: int, gain: float = 1.0) -> np.ndarray:
    pcm = np.empty(frames * channels * bits // 8, dtype=np.uint8)
    rc = lib().oracle_synth_pcm(seed, first_frame, frames, channels, bits, int(big_endian), rate,
                                gain, _b(pcm))
    if rc:
        raise ValueError("unsupported bit depth")
    return pcm


def process(pcm: np.ndarray, frames: int, channels: int, bits: int, big_endian: bool,
            fc_norm: float, bw_norm: float, normalize: bool, halo_l: int = 0, halo_r: int = 0,
            scale_in: float = 0.0):
    """Whole hot path in "hi" mode.  Returns dict(pcm, y, peak, scale)."""
    pcm = np.ascontiguousarray(pcm, dtype=np.uint8)
    nb = bits // 8
    assert pcm.size == (halo_l + frames + halo_r) * channels * nb
    out = np.empty(frames * channels * nb, dtype=np.uint8)
    y = np.empty((channels, frames), dtype=np.float64)
    pk = C.c_double()
    sc = C.c_double()
    rc = lib().oracle_process(_b(pcm), frames, channels, bits, int(big_endian), halo_l, halo_r,
                              fc_norm, bw_norm, int(normalize), scale_in, _b(out), _d(y),
                              C.byref(pk), C.byref(sc))
    if rc:
        raise ValueError(f"oracle_process rc={rc}")
    return {"pcm": out, "y": y, "peak": pk.value, "scale": sc.value}
