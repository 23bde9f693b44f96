"""CPU tests of the boundary: the C-ABI library loads, exports every symbol the
header declares, and refuses to compute without a device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def header_symbols(name="fir_gpu.h"):
    text = open(os.path.join(ROOT, "include", name)).read()
    return re.findall(r"FIR_GPU_API\s+[\w\s\*]+?\b(fir_gpu_\w+)\s*\(", text)


def test_header_and_binding_agree():
    from audio_fir_filter_b200 import capi

    syms = header_symbols()
    assert len(syms) >= 25
    assert sorted(syms) == sorted(capi.SYMBOLS)
    assert sorted(header_symbols("fir_gpu_dev.h")) == sorted(capi.DEV_SYMBOLS)
    assert not set(capi.SYMBOLS) & set(capi.DEV_SYMBOLS)


def test_boundary_header_carries_no_measurement_or_test_entry_points():
    """Probes, synthetic input, kernel variants and test hooks live in fir_gpu_dev.h, which the
    C++ host never includes."""
    pub = header_symbols()
    for s in ("fir_gpu_fp64_peak", "fir_gpu_set_variant", "fir_gpu_synth_pcm_dev", "fir_gpu_set_x_budget",
              "fir_gpu_kernel_from_taps", "fir_gpu_copy_probe", "fir_gpu_test_fail_next_create"):
        assert s not in pub, s
    for f in os.listdir(os.path.join(ROOT, "host")):
        if f.endswith((".cpp", ".hpp")):
            assert "fir_gpu_dev.h" not in open(os.path.join(ROOT, "host", f)).read(), f


def test_library_exports_every_declared_symbol():
    from audio_fir_filter_b200 import capi

    assert os.path.exists(capi.LIB_PATH), "libfir_gpu.so was not built (run __graft_entry__.build())"
    L = ctypes.CDLL(capi.LIB_PATH)
    for s in header_symbols() + header_symbols("fir_gpu_dev.h"):
        assert hasattr(L, s), s
    assert capi.lib() is not None


def test_no_cpu_fallback_without_a_device():
    import torch

    from audio_fir_filter_b200 import capi

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert capi.device_count() == 0
    with pytest.raises(capi.FirGpuError) as e:
        capi.Context(0)
    assert e.value.code == capi.ERR_NO_DEVICE
    assert "no CPU path" in str(e.value) or "no CUDA device" in str(e.value)


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under the package or host/ may name it."""
    bad = []
    for base in ("audio_fir_filter_b200", "host", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp", "Makefile")):
                    t = open(os.path.join(dp, f), errors="replace").read()
                    if re.search(r"^\s*(import|from)\s+oracle\b|liboracle|fir_oracle|oracle/", t, re.M):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_pcm_struct_layout_matches_header():
    from audio_fir_filter_b200 import capi

    assert ctypes.sizeof(capi.PcmFormat) == 40
    assert capi.PcmFormat.halo_left.offset == 24
    assert ctypes.sizeof(capi.Timing) == 64
