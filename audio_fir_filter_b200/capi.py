"""ctypes binding of ``include/fir_gpu.h`` (the C-ABI in ``libfir_gpu.so``).

This is the same binding a maintainer of the reference would add around
``process_file``; it is used by the tests and by
``bench.py``.  There is no CPU path: without the built CUDA library, or without
an sm_100 device, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
# FIR_GPU_LIB selects another build of the same library (tools/sweep_variants.py uses the
# -DFIR_ALL_VARIANTS build, libfir_gpu_sweep.so); there is still no fallback of any kind.
LIB_PATH = os.environ.get("FIR_GPU_LIB") or os.path.join(_PKG, "libfir_gpu.so")

OK, ERR_NO_DEVICE, ERR_INVALID, ERR_CUDA, ERR_STATE, ERR_NOMEM = range(6)
_CODE_NAMES = {1: "NO_DEVICE", 2: "INVALID", 3: "CUDA", 4: "STATE", 5: "NOMEM"}


class FirGpuError(RuntimeError):
    """A non-zero status from the C-ABI.  ``code`` is the FIR_GPU_ERR_* value."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"fir_gpu error {_CODE_NAMES.get(code, code)}: {msg}")
        self.code = code
        self.message = msg


class PcmFormat(C.Structure):
    """``struct fir_gpu_pcm``."""

    _fields_ = [
        ("frames", C.c_int64),
        ("channels", C.c_int32),
        ("bits", C.c_int32),
        ("big_endian", C.c_int32),
        ("halo_left", C.c_int64),
        ("halo_right", C.c_int64),
    ]


class Timing(C.Structure):
    """``struct fir_gpu_timing`` (milliseconds, CUDA events on the context's stream)."""

    _fields_ = [
        ("h2d_ms", C.c_double),
        ("decode_ms", C.c_double),
        ("fir_ms", C.c_double),
        ("peak_ms", C.c_double),
        ("encode_ms", C.c_double),
        ("d2h_ms", C.c_double),
        ("fir_launches", C.c_int64),
        ("other_launches", C.c_int64),
    ]

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_}


# Every symbol include/fir_gpu.h (the drop-in boundary) declares: name -> (restype, argtypes).
_vp = C.c_void_p
_i64 = C.c_int64
_dp = C.POINTER(C.c_double)
SYMBOLS = {
    "fir_gpu_device_count": (C.c_int, []),
    "fir_gpu_create": (C.c_int, [C.c_int, C.POINTER(_vp)]),
    "fir_gpu_destroy": (None, [_vp]),
    "fir_gpu_last_error": (C.c_char_p, []),
    "fir_gpu_set_stream": (C.c_int, [_vp, _vp]),
    "fir_gpu_synchronize": (C.c_int, [_vp]),
    "fir_gpu_host_alloc": (_vp, [C.c_size_t]),
    "fir_gpu_host_free": (None, [_vp]),
    "fir_gpu_build_kernel": (C.c_int, [_vp, C.c_double, C.c_double, C.POINTER(_vp), C.POINTER(_i64)]),
    "fir_gpu_kernel_num_taps": (_i64, [_vp]),
    "fir_gpu_kernel_taps": (C.c_int, [_vp, _vp, _dp, _i64]),
    "fir_gpu_kernel_free": (None, [_vp]),
    "fir_gpu_apply": (C.c_int, [_vp, _vp, _vp, C.POINTER(PcmFormat)]),
    "fir_gpu_apply_dev": (C.c_int, [_vp, _vp, _vp, C.POINTER(PcmFormat)]),
    "fir_gpu_apply_begin": (C.c_int, [_vp, _vp, C.POINTER(PcmFormat)]),
    "fir_gpu_apply_feed": (C.c_int, [_vp, _vp, C.c_size_t]),
    "fir_gpu_apply_end": (C.c_int, [_vp]),
    "fir_gpu_set_progress": (C.c_int, [_vp, _vp, _vp]),
    "fir_gpu_encode_range": (C.c_int, [_vp, C.c_double, _i64, _i64, _vp]),
    "fir_gpu_process": (C.c_int, [_vp, _vp, _vp, C.POINTER(PcmFormat), C.c_int, _vp, _dp, _dp]),
    "fir_gpu_filter_f64": (C.c_int, [_vp, _vp, _dp, _i64, C.c_int32, _dp]),
    "fir_gpu_parked": (C.c_int, [_vp, _dp, _i64, C.c_int32]),
    "fir_gpu_parked_range": (C.c_int, [_vp, _dp, _i64, _i64, C.c_int32]),
    "fir_gpu_peak": (C.c_int, [_vp, _dp]),
    "fir_gpu_peak_dev": (C.c_int, [_vp, C.POINTER(_vp)]),
    "fir_gpu_peak_recompute": (C.c_int, [_vp, _dp]),
    "fir_gpu_comm_prepare": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "fir_gpu_allreduce_peak": (C.c_int, [C.POINTER(_vp), C.c_int, _dp]),
    "fir_gpu_encode": (C.c_int, [_vp, C.c_double, _vp]),
    "fir_gpu_encode_dev": (C.c_int, [_vp, C.c_double, _vp]),
    "fir_gpu_last_timing": (C.c_int, [_vp, C.POINTER(Timing)]),
}
# include/fir_gpu_dev.h: measurement, synthetic input, tuning and test hooks (bench.py, tools/, tests/).
DEV_SYMBOLS = {
    "fir_gpu_kernel_from_taps": (C.c_int, [_vp, _dp, _i64, C.POINTER(_vp)]),
    "fir_gpu_synth_pcm_dev": (C.c_int, [_vp, C.c_uint64, _i64, _i64, C.c_int32, C.c_int32, C.c_int32,
                                        _i64, C.c_double, _vp]),
    "fir_gpu_fp64_peak": (C.c_int, [_vp, C.c_int, C.c_double, _dp]),
    "fir_gpu_copy_probe": (C.c_int, [_vp, _vp, C.c_size_t, C.c_int, _dp]),
    "fir_gpu_reserve": (C.c_int, [_vp, _vp, C.POINTER(PcmFormat), C.c_int]),
    "fir_gpu_set_variant": (C.c_int, [_vp, C.c_int]),
    "fir_gpu_variant_count": (C.c_int, []),
    "fir_gpu_variant_name": (C.c_char_p, [C.c_int]),
    "fir_gpu_set_codec_geometry": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int]),
    "fir_gpu_test_fail_next_create": (C.c_int, [C.c_int]),
    "fir_gpu_set_x_budget": (C.c_int, [_vp, _i64]),
}

_lib = None


def build(force: bool = False) -> str:
    """Compile libfir_gpu.so for sm_100a with nvcc (in-tree)."""
    csrc = os.path.join(_PKG, "csrc")
    args = ["make", "-C", csrc]
    if force:
        args.append("-B")
    r = subprocess.run(args, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building libfir_gpu.so failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


def lib() -> C.CDLL:
    """The loaded C-ABI library.  Raises if it was never built -- there is no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FirGpuError(ERR_NO_DEVICE, f"{LIB_PATH} is missing: build it with "
                              "`make -C audio_fir_filter_b200/csrc` (no CPU fallback exists)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in {**SYMBOLS, **DEV_SYMBOLS}.items():
            f = getattr(L, name)  # AttributeError if the header and library disagree
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def _check(rc: int) -> None:
    if rc != OK:
        raise FirGpuError(rc, lib().fir_gpu_last_error().decode("utf-8", "replace"))


def device_count() -> int:
    return int(lib().fir_gpu_device_count())


def _ptr(buf) -> int:
    """Host/device address of a numpy array, torch tensor or plain int."""
    if isinstance(buf, int):
        return buf
    if isinstance(buf, np.ndarray):
        return buf.ctypes.data
    if hasattr(buf, "data_ptr"):
        return int(buf.data_ptr())
    raise TypeError(f"cannot take the address of {type(buf)!r}")


class Kernel:
    """A tap array resident in HBM (``fir_gpu_kernel``)."""

    def __init__(self, ctx: "Context", handle: int, half_len: int):
        self._ctx = ctx
        self._h = handle
        self.half_len = half_len

    @property
    def num_taps(self) -> int:
        return int(lib().fir_gpu_kernel_num_taps(self._h))

    def taps(self) -> np.ndarray:
        n = self.num_taps
        out = np.empty(n, dtype=np.float64)
        _check(lib().fir_gpu_kernel_taps(self._ctx._h, self._h, out.ctypes.data_as(_dp), n))
        return out

    def free(self) -> None:
        if self._h:
            lib().fir_gpu_kernel_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One per GPU (``fir_gpu_ctx``); one host thread at a time."""

    def __init__(self, device: int = 0):
        h = _vp()
        _check(lib().fir_gpu_create(device, C.byref(h)))
        self._h = h.value
        self.device = device

    def close(self) -> None:
        if self._h:
            lib().fir_gpu_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- stream --------------------------------------------------------------
    def set_stream(self, cuda_stream: int | None) -> None:
        _check(lib().fir_gpu_set_stream(self._h, cuda_stream or None))

    def synchronize(self) -> None:
        _check(lib().fir_gpu_synchronize(self._h))

    # ---- kernel --------------------------------------------------------------
    def build_kernel(self, fc_norm: float, bw_norm: float) -> Kernel:
        k = _vp()
        half = _i64()
        _check(lib().fir_gpu_build_kernel(self._h, fc_norm, bw_norm, C.byref(k), C.byref(half)))
        return Kernel(self, k.value, int(half.value))

    def kernel_from_taps(self, taps: np.ndarray) -> Kernel:
        taps = np.ascontiguousarray(taps, dtype=np.float64)
        k = _vp()
        _check(lib().fir_gpu_kernel_from_taps(self._h, taps.ctypes.data_as(_dp), taps.size, C.byref(k)))
        return Kernel(self, k.value, (taps.size - 1) // 2)

    # ---- apply / peak / encode -------------------------------------------------
    @staticmethod
    def _fmt(frames, channels, bits, big_endian, halo_left=0, halo_right=0) -> PcmFormat:
        return PcmFormat(int(frames), int(channels), int(bits), int(bool(big_endian)), int(halo_left),
                         int(halo_right))

    def apply(self, kernel: Kernel, pcm_host, frames: int, channels: int, bits: int, big_endian: bool,
              halo_left: int = 0, halo_right: int = 0) -> None:
        fmt = self._fmt(frames, channels, bits, big_endian, halo_left, halo_right)
        if isinstance(pcm_host, np.ndarray):
            need = (halo_left + frames + halo_right) * channels * bits // 8
            if pcm_host.nbytes < need:
                raise ValueError(f"PCM buffer holds {pcm_host.nbytes} bytes, format needs {need}")
        _check(lib().fir_gpu_apply(self._h, kernel._h, _ptr(pcm_host) or None, C.byref(fmt)))

    def apply_dev(self, kernel: Kernel, pcm_dev, frames: int, channels: int, bits: int, big_endian: bool,
                  halo_left: int = 0, halo_right: int = 0) -> None:
        fmt = self._fmt(frames, channels, bits, big_endian, halo_left, halo_right)
        _check(lib().fir_gpu_apply_dev(self._h, kernel._h, _ptr(pcm_dev) or None, C.byref(fmt)))

    def apply_streamed(self, kernel: Kernel, pieces, frames: int, channels: int, bits: int, big_endian: bool,
                       halo_left: int = 0, halo_right: int = 0) -> None:
        """fir_gpu_apply_begin / _feed / _end: ``pieces`` is an iterable of host byte buffers that,
        concatenated, are the buffer fir_gpu_apply would receive."""
        fmt = self._fmt(frames, channels, bits, big_endian, halo_left, halo_right)
        _check(lib().fir_gpu_apply_begin(self._h, kernel._h, C.byref(fmt)))
        keep = []
        for p in pieces:
            keep.append(p)          # keep the last two alive while their copies may be in flight
            keep = keep[-2:]
            n = p.nbytes if isinstance(p, np.ndarray) else len(p)
            _check(lib().fir_gpu_apply_feed(self._h, _ptr(p) if n else None, n))
        _check(lib().fir_gpu_apply_end(self._h))

    def process(self, kernel: Kernel, pcm_in, frames: int, channels: int, bits: int, big_endian: bool,
                normalize: bool, pcm_out, halo_left: int = 0, halo_right: int = 0) -> tuple[float, float]:
        """fir_gpu_process: the whole path of one host payload in one call -> (peak, scale)."""
        fmt = self._fmt(frames, channels, bits, big_endian, halo_left, halo_right)
        pk, sc = C.c_double(), C.c_double()
        _check(lib().fir_gpu_process(self._h, kernel._h, _ptr(pcm_in) or None, C.byref(fmt), int(bool(normalize)),
                                     _ptr(pcm_out) or None, C.byref(pk), C.byref(sc)))
        return float(pk.value), float(sc.value)

    def set_progress(self, fn) -> None:
        """fn(done_frames, total_frames) or None.  Called from a CUDA callback thread."""
        if fn is None:
            self._progress_cb = None
            _check(lib().fir_gpu_set_progress(self._h, None, None))
            return
        proto = C.CFUNCTYPE(None, C.c_int64, C.c_int64, C.c_void_p)
        self._progress_cb = proto(lambda d, t, u: fn(d, t))
        _check(lib().fir_gpu_set_progress(self._h, C.cast(self._progress_cb, _vp), None))

    def encode_range(self, scale: float, first_frame: int, frames: int, out) -> None:
        _check(lib().fir_gpu_encode_range(self._h, float(scale), first_frame, frames, _ptr(out) or None))

    def filter_f64(self, kernel: Kernel, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.ndim == 1:
            x = x[None, :]
        ch, frames = x.shape
        y = np.empty_like(x)
        _check(lib().fir_gpu_filter_f64(self._h, kernel._h, x.ctypes.data_as(_dp), frames, ch,
                                        y.ctypes.data_as(_dp)))
        return y

    def parked(self, frames: int, channels: int) -> np.ndarray:
        y = np.empty((channels, frames), dtype=np.float64)
        _check(lib().fir_gpu_parked(self._h, y.ctypes.data_as(_dp), frames, channels))
        return y

    def parked_range(self, first_frame: int, frames: int, channels: int) -> np.ndarray:
        y = np.empty((channels, frames), dtype=np.float64)
        _check(lib().fir_gpu_parked_range(self._h, y.ctypes.data_as(_dp), first_frame, frames, channels))
        return y

    def peak(self) -> float:
        v = C.c_double()
        _check(lib().fir_gpu_peak(self._h, C.byref(v)))
        return float(v.value)

    def peak_dev(self) -> int:
        p = _vp()
        _check(lib().fir_gpu_peak_dev(self._h, C.byref(p)))
        return int(p.value)

    def peak_recompute(self) -> float:
        v = C.c_double()
        _check(lib().fir_gpu_peak_recompute(self._h, C.byref(v)))
        return float(v.value)

    def encode(self, scale: float, out) -> None:
        _check(lib().fir_gpu_encode(self._h, float(scale), _ptr(out) or None))

    def encode_dev(self, scale: float, out_dev) -> None:
        _check(lib().fir_gpu_encode_dev(self._h, float(scale), _ptr(out_dev) or None))

    # ---- measurement -----------------------------------------------------------
    def last_timing(self) -> dict:
        t = Timing()
        _check(lib().fir_gpu_last_timing(self._h, C.byref(t)))
        return t.as_dict()

    def synth_pcm_dev(self, seed: int, first_frame: int, frames: int, channels: int, bits: int,
                      big_endian: bool, rate: int, gain: float, pcm_dev) -> None:
        _check(lib().fir_gpu_synth_pcm_dev(self._h, seed, first_frame, frames, channels, bits,
                                           int(bool(big_endian)), rate, gain, _ptr(pcm_dev)))

    def fp64_peak(self, kind: int, seconds: float = 0.2) -> float:
        v = C.c_double()
        _check(lib().fir_gpu_fp64_peak(self._h, kind, seconds, C.byref(v)))
        return float(v.value)

    def copy_probe(self, host_buf, nbytes: int, direction: int) -> float:
        """One plain pinned copy of ``nbytes`` (0 = H2D, 1 = D2H) on the context's stream -> milliseconds."""
        v = C.c_double()
        _check(lib().fir_gpu_copy_probe(self._h, _ptr(host_buf), nbytes, direction, C.byref(v)))
        return float(v.value)

    def reserve(self, kernel: Kernel, frames: int, channels: int, bits: int, big_endian: bool, halo_left: int = 0,
                halo_right: int = 0, host_path: bool = False) -> None:
        fmt = self._fmt(frames, channels, bits, big_endian, halo_left, halo_right)
        _check(lib().fir_gpu_reserve(self._h, kernel._h, C.byref(fmt), int(host_path)))

    def set_codec_geometry(self, tile_bytes: int, threads: int, carveout_pct: int = 50) -> None:
        _check(lib().fir_gpu_set_codec_geometry(self._h, tile_bytes, threads, carveout_pct))

    def set_variant(self, variant: int) -> None:
        _check(lib().fir_gpu_set_variant(self._h, variant))

    def set_x_budget(self, nbytes: int) -> None:
        _check(lib().fir_gpu_set_x_budget(self._h, nbytes))


def allreduce_peak(ctxs: list[Context]) -> float:
    """fir_gpu_allreduce_peak: ONE ncclAllReduce(max) over the peak scalars of contexts on distinct
    devices of this process (sample-block mode); every context then holds the global peak."""
    arr = (_vp * len(ctxs))(*[c._h for c in ctxs])
    v = C.c_double()
    _check(lib().fir_gpu_allreduce_peak(arr, len(ctxs), C.byref(v)))
    return float(v.value)


def comm_prepare(ctxs: list[Context]) -> None:
    arr = (_vp * len(ctxs))(*[c._h for c in ctxs])
    _check(lib().fir_gpu_comm_prepare(arr, len(ctxs)))


def variant_names() -> list[str]:
    L = lib()
    return [L.fir_gpu_variant_name(i).decode() for i in range(L.fir_gpu_variant_count())]


class PinnedBuffer:
    """Page-locked host bytes (``fir_gpu_host_alloc``) viewed as a numpy uint8 array."""

    def __init__(self, nbytes: int):
        p = lib().fir_gpu_host_alloc(nbytes)
        if not p:
            raise FirGpuError(ERR_NOMEM, lib().fir_gpu_last_error().decode())
        self._p = p
        self.nbytes = nbytes
        self.array = np.ctypeslib.as_array((C.c_uint8 * max(nbytes, 1)).from_address(p))[:nbytes]

    def free(self) -> None:
        if self._p:
            self.array = None
            lib().fir_gpu_host_free(self._p)
            self._p = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
