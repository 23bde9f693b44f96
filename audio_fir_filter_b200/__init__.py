"""B200-native (sm_100a) low-cut FIR hot path of diskerror/audio-fir-filter.

The product is ``libfir_gpu.so`` (hand-written CUDA behind the C-ABI in
``include/fir_gpu.h``) and the C++23 host in ``host/``.  This package holds the
kernels' sources (``csrc/``), the ctypes binding of the C-ABI (``capi``) and a
Python mirror of the host's call sequence (``process``) for tests and bench.
"""
from . import capi  # noqa: F401
from .capi import Context, FirGpuError, Kernel, PinnedBuffer, device_count  # noqa: F401
from .process import (Block, FilterOptions, PcmInfo, plan_blocks, process_pcm,  # noqa: F401
                      process_pcm_sharded, scale_for_peak)
