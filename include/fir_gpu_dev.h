/*
 * fir_gpu_dev.h -- measurement, synthetic-input and tuning entry points of
 * libfir_gpu.so.  NOT part of the drop-in boundary (that is fir_gpu.h): nothing
 * a host of the reference's process_file() needs is declared here.  Users:
 * bench.py, tools/ and tests/ only -- host/ never includes this header.
 */
#ifndef FIR_GPU_DEV_H
#define FIR_GPU_DEV_H

#include "fir_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

/* A kernel from caller-supplied taps (n_taps odd): for parity tests that must
 * feed the device the oracle's exact taps. */
FIR_GPU_API int fir_gpu_kernel_from_taps(fir_gpu_ctx *ctx, const double *taps, int64_t n_taps,
                                         fir_gpu_kernel **out);

/* Counter-based synthetic PCM (same integers as oracle_synth_pcm), written to
 * device memory: any window of any config without materialising the file. */
FIR_GPU_API int fir_gpu_synth_pcm_dev(fir_gpu_ctx *ctx, uint64_t seed, int64_t first_frame,
                                      int64_t frames, int32_t channels, int32_t bits,
                                      int32_t big_endian, int64_t rate, double gain, void *pcm_dev);

/* Register-resident FP64 throughput probes (TFLOP/s) used as the measured
 * roofline denominator: kind 0 = DFMA pipe, 1 = DMMA (mma.sync m8n8k4 f64). */
FIR_GPU_API int fir_gpu_fp64_peak(fir_gpu_ctx *ctx, int kind, double seconds, double *tflops);

/* One plain cudaMemcpyAsync of `bytes` between host_buf (pinned or pageable) and the
 * context's PCM staging buffer, timed with CUDA events on the context's stream:
 * dir 0 = host->device, 1 = device->host.  The PCIe ceiling the end-to-end numbers
 * are compared with (N ranks call it at the same moment). */
FIR_GPU_API int fir_gpu_copy_probe(fir_gpu_ctx *ctx, void *host_buf, size_t bytes, int dir, double *ms);

/* Allocate now every device buffer an apply of this format with this kernel will need
 * (host_path != 0: the staging buffers of the host-buffer entry points too), so that a
 * single timed pass does not measure cudaMalloc of tens of gigabytes. */
FIR_GPU_API int fir_gpu_reserve(fir_gpu_ctx *ctx, const fir_gpu_kernel *k, const fir_gpu_pcm *fmt, int host_path);

/* FIR kernel variant (0 = default).  The product build carries the default DMMA
 * kernel and one DFMA comparison kernel; -DFIR_ALL_VARIANTS adds the shapes of the
 * tuning sweeps (tools/sweep_variants.py). */
FIR_GPU_API int fir_gpu_set_variant(fir_gpu_ctx *ctx, int variant);
FIR_GPU_API int fir_gpu_variant_count(void);
FIR_GPU_API const char *fir_gpu_variant_name(int variant);

/* Tile size (interleaved bytes staged per trip; 0 = the default of 4096 samples), threads per CTA (128 or 256) and shared-memory
 * carveout (percent of the SM's array, -1 = the driver's choice) of the PCM decode / encode
 * kernels, for tuning sweeps; the result does not depend on them. */
FIR_GPU_API int fir_gpu_set_codec_geometry(fir_gpu_ctx *ctx, int tile_bytes, int threads, int smem_carveout_pct);

/* Test hook: the next n fir_gpu_create calls fail after their streams, events and device
 * buffers exist (proves that a half-built context is torn down completely). */
FIR_GPU_API int fir_gpu_test_fail_next_create(int n);

/* Bound on the decoded FP64 input scratch (files longer than this stream through
 * it chunk by chunk; the result does not depend on it). */
FIR_GPU_API int fir_gpu_set_x_budget(fir_gpu_ctx *ctx, int64_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* FIR_GPU_DEV_H */
