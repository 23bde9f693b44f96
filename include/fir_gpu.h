/*
 * fir_gpu.h -- C-ABI of the B200 (sm_100a) low-cut FIR hot path.
 *
 * This is the drop-in boundary for the arithmetic half of the reference's
 * process_file() (diskerror/audio-fir-filter, ProcessFile.cp:27-120).  The
 * reference has no FFI of its own; the seam is cut where process_file hands
 * work to the un-vendored c_lib, and every entry point below names the
 * reference call site it replaces.  A host that keeps the reference's CLI and
 * chunk-preserving file I/O calls, per file:
 *
 *     fir_gpu_build_kernel   <- WindowedSinc<float64_t>(fc,bw) + makeLowCut()   ProcessFile.cp:48-50
 *     fir_gpu_apply          <- readAll() + the per-channel thread fan-out over  ProcessFile.cp:41,57-87
 *                               apply_filter_range()                             FilterCore.h:20-79
 *     fir_gpu_peak           <- the max_mag() loop                               ProcessFile.cp:92-96
 *     [host: scale rule  maxMag > 1 || normalize                                 ProcessFile.cp:98]
 *     fir_gpu_encode         <- AudioSamples::normalize + writeAll(buf, true)    ProcessFile.cp:100,117
 *   or all of it in one call:
 *     fir_gpu_process        <- ProcessFile.cp:41-101,117
 *   or piece by piece (file reads / writes overlapping the GPU):
 *     fir_gpu_apply_begin / _feed / _end, fir_gpu_encode_range, fir_gpu_set_progress
 *
 * Plain C: opaque handles, plain pointers and sizes, int status codes (0 = ok),
 * no exceptions across the boundary, no CPU fallback -- without a usable
 * sm_100 device every call fails with FIR_GPU_ERR_NO_DEVICE.
 *
 * Threading: one context per GPU, used by one host thread at a time (the
 * reference's -t fan-out is replaced by the device grid).  All work of a context
 * is ordered on one CUDA stream (its own, or the one given to
 * fir_gpu_set_stream).
 */
#ifndef FIR_GPU_H
#define FIR_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define FIR_GPU_API __attribute__((visibility("default")))
#else
#define FIR_GPU_API
#endif

typedef struct fir_gpu_ctx fir_gpu_ctx;       /* per-device state: stream, parked signal, scratch */
typedef struct fir_gpu_kernel fir_gpu_kernel; /* one tap array resident in HBM */

enum {
	FIR_GPU_OK = 0,
	FIR_GPU_ERR_NO_DEVICE = 1, /* no CUDA device / not sm_100 / driver missing */
	FIR_GPU_ERR_INVALID = 2,   /* bad argument */
	FIR_GPU_ERR_CUDA = 3,      /* a CUDA call failed; see fir_gpu_last_error() */
	FIR_GPU_ERR_STATE = 4,     /* call order: nothing parked, wrong context, ... */
	FIR_GPU_ERR_NOMEM = 5
};

/* PCM layout of one block of interleaved frames (what AudioFormat tells the
 * reference: channels(), bits, endianness; ProcessFile.cp:35,43). */
typedef struct fir_gpu_pcm {
	int64_t frames;     /* frames to PRODUCE (the block this call owns) */
	int32_t channels;   /* >= 1 */
	int32_t bits;       /* 16, 24 or 32, signed two's complement */
	int32_t big_endian; /* 0 = WAVE order, 1 = AIFF order */
	/* Sample-block sharding (single long file across GPUs): real frames that
	 * precede / follow the block in the buffer handed to fir_gpu_apply.  Each is
	 * at most half_len (more is accepted and ignored); whatever is missing is
	 * implicit zero, i.e. a true file edge (FilterCore.h:57-61,72-76). */
	int64_t halo_left;
	int64_t halo_right;
} fir_gpu_pcm;

/* Per-phase device times of the last apply/encode on this context, measured
 * with CUDA events on the context's stream (milliseconds). */
typedef struct fir_gpu_timing {
	double h2d_ms, decode_ms, fir_ms, peak_ms, encode_ms, d2h_ms;
	int64_t fir_launches, other_launches; /* kernels launched by the last apply + encode */
} fir_gpu_timing;

/* ---- life cycle ---------------------------------------------------------- */

/* Number of usable sm_100 devices (0 when there is none; never a CPU path). */
FIR_GPU_API int fir_gpu_device_count(void);

FIR_GPU_API int fir_gpu_create(int device, fir_gpu_ctx **out);
FIR_GPU_API void fir_gpu_destroy(fir_gpu_ctx *ctx);

/* Message of the last failure on the calling thread ("" if none). */
FIR_GPU_API const char *fir_gpu_last_error(void);

/* Order this context's work on an existing cudaStream_t (e.g. the caller's
 * framework stream).  NULL restores the context's own stream. */
FIR_GPU_API int fir_gpu_set_stream(fir_gpu_ctx *ctx, void *cuda_stream);
FIR_GPU_API int fir_gpu_synchronize(fir_gpu_ctx *ctx);

/* Pinned host memory for the PCM payload buffers (optional; pageable works). */
FIR_GPU_API void *fir_gpu_host_alloc(size_t bytes);
FIR_GPU_API void fir_gpu_host_free(void *p);

/* ---- fir_gpu_build_kernel  (ProcessFile.cp:48-50) ------------------------ */

/* Blackman windowed-sinc low-pass of order M = round(4/bw) forced even,
 * normalised to unity DC gain with a double-double sum, then spectrally
 * inverted to the low-cut.  fc_norm = freq / sampleRate and bw_norm = slope /
 * sampleRate, exactly the two arguments at ProcessFile.cp:49.  *half_len
 * receives M/2 (WindowedSinc::getMo2(), FilterCore.h:29). */
FIR_GPU_API int fir_gpu_build_kernel(fir_gpu_ctx *ctx, double fc_norm, double bw_norm,
                                     fir_gpu_kernel **out, int64_t *half_len);

FIR_GPU_API int64_t fir_gpu_kernel_num_taps(const fir_gpu_kernel *k);
/* Copy the M+1 taps back to the host (parity checks). */
FIR_GPU_API int fir_gpu_kernel_taps(fir_gpu_ctx *ctx, const fir_gpu_kernel *k, double *taps_out,
                                    int64_t n);
FIR_GPU_API void fir_gpu_kernel_free(fir_gpu_kernel *k);

/* ---- fir_gpu_apply  (ProcessFile.cp:41,57-87 / FilterCore.h:20-79) ------- */

/* Filter phase.  pcm_host points at the first byte of frame (-halo_left): the
 * buffer holds halo_left + frames + halo_right interleaved frames.  Uploads,
 * decodes to planar FP64, runs
 *     y[n] = sum_{k=0..M} h[k] * x[n - M/2 + k],  x = 0 outside the file,
 * for every channel and n in [0, frames), and PARKS y (FP64) in HBM together
 * with its peak magnitude.  Asynchronous on the context's stream. */
FIR_GPU_API int fir_gpu_apply(fir_gpu_ctx *ctx, const fir_gpu_kernel *k, const void *pcm_host,
                              const fir_gpu_pcm *fmt);

/* Streamed filter phase: the same work as fir_gpu_apply with the payload arriving
 * piece by piece, so that file reads overlap the upload and the FIR (the reference
 * reads the whole file first, ProcessFile.cp:41).  begin announces the layout; each
 * feed hands over the NEXT `bytes` of the buffer fir_gpu_apply would have received
 * (any split, in order) and starts every chunk whose samples have landed; end starts
 * the rest and parks the signal.  When feed(n) returns, the host buffer given to
 * feed(n-1) is free again (alternate two buffers); after end all are. */
FIR_GPU_API int fir_gpu_apply_begin(fir_gpu_ctx *ctx, const fir_gpu_kernel *k, const fir_gpu_pcm *fmt);
FIR_GPU_API int fir_gpu_apply_feed(fir_gpu_ctx *ctx, const void *pcm_host, size_t bytes);
FIR_GPU_API int fir_gpu_apply_end(fir_gpu_ctx *ctx);

/* Progress hook (the reference reports progress from apply_filter_range,
 * FilterCore.h:38-54, ProgressBar.h:58-82): fn(done_frames, total_frames, user) is
 * called as each chunk of the FIR COMPLETES on the device, from a CUDA callback
 * thread -- it must not call into this library or CUDA.  NULL removes it. */
typedef void (*fir_gpu_progress_fn)(int64_t done_frames, int64_t total_frames, void *user);
FIR_GPU_API int fir_gpu_set_progress(fir_gpu_ctx *ctx, fir_gpu_progress_fn fn, void *user);

/* Same with the PCM already resident in device memory (no H2D). */
FIR_GPU_API int fir_gpu_apply_dev(fir_gpu_ctx *ctx, const fir_gpu_kernel *k, const void *pcm_dev,
                                  const fir_gpu_pcm *fmt);

/* The bare FIR on caller-supplied planar FP64 host samples x[channels][frames]
 * -> y[channels][frames] (host), bypassing the PCM codec.  This is the
 * apply_filter_range() equivalent the 1e-12 parity tests call. */
FIR_GPU_API int fir_gpu_filter_f64(fir_gpu_ctx *ctx, const fir_gpu_kernel *k, const double *x_host,
                                   int64_t frames, int32_t channels, double *y_host);

/* Copy the parked FP64 signal, planar [channels][frames], to the host. */
FIR_GPU_API int fir_gpu_parked(fir_gpu_ctx *ctx, double *y_host, int64_t frames, int32_t channels);

/* A window of it: frames [first_frame, first_frame + frames) of every channel,
 * planar [channels][frames] (spot checks of files too large to copy back whole). */
FIR_GPU_API int fir_gpu_parked_range(fir_gpu_ctx *ctx, double *y_host, int64_t first_frame, int64_t frames,
                                     int32_t channels);

/* ---- fir_gpu_peak  (ProcessFile.cp:92-96) -------------------------------- */

/* max over channels and frames of |y| of the parked signal on THIS device.
 * Synchronises the stream.  In sample-block mode the caller max-reduces the
 * per-device values (NCCL allreduce-max) before choosing the scale. */
FIR_GPU_API int fir_gpu_peak(fir_gpu_ctx *ctx, double *peak);
/* Device address of that FP64 scalar (valid until the next apply), so the
 * all-reduce can run on it without a host round trip. */
FIR_GPU_API int fir_gpu_peak_dev(fir_gpu_ctx *ctx, void **peak_dev);
/* Recompute the peak from the parked signal with the stand-alone warp-reduced
 * kernel instead of the value fused into the FIR epilogue. */
FIR_GPU_API int fir_gpu_peak_recompute(fir_gpu_ctx *ctx, double *peak);

/* ---- the peak over sample blocks on several GPUs  (ProcessFile.cp:92-96) -- */

/* Sample-block mode: one long file, one contiguous block per device, every block
 * filtered on its own context of THIS process.  The reference's peak is a max over the
 * whole file, so the per-device peaks are max-reduced with ONE ncclAllReduce(ncclMax) of
 * the 8-byte device scalars over NVLink, in place and ordered on each context's stream;
 * afterwards every context holds the global peak (fir_gpu_peak / fir_gpu_peak_dev return
 * it) and *peak receives it.  This is the only collective of the whole path.
 * The communicator of a device set is created on first use and kept;
 * fir_gpu_comm_prepare creates it ahead of time (from another host thread, while the
 * blocks are still being filtered).  NCCL (libnccl.so.2) is loaded on first use:
 * FIR_GPU_ERR_STATE when it cannot be -- the caller may then take the max of the
 * fir_gpu_peak values itself (n doubles; nothing is computed on the CPU either way).
 * One-process-per-GPU hosts (torch.distributed, MPI) run their own all-reduce on
 * fir_gpu_peak_dev instead. */
FIR_GPU_API int fir_gpu_comm_prepare(fir_gpu_ctx *const *ctxs, int n);
FIR_GPU_API int fir_gpu_allreduce_peak(fir_gpu_ctx *const *ctxs, int n, double *peak);

/* ---- fir_gpu_encode  (ProcessFile.cp:100,117) ---------------------------- */

/* Encode phase: q = clamp(rint(y * scale * 2^(bits-1))) (ties to even, clamp to
 * the signed range, no dither), interleave, endian as in the matching apply;
 * writes frames*channels*bits/8 bytes to pcm_host.  Synchronous on return. */
FIR_GPU_API int fir_gpu_encode(fir_gpu_ctx *ctx, double scale, void *pcm_host);
/* Frames [first_frame, first_frame + frames) only, into pcm_host (that many frames):
 * lets the host write the output file piece by piece while the next piece encodes. */
FIR_GPU_API int fir_gpu_encode_range(fir_gpu_ctx *ctx, double scale, int64_t first_frame, int64_t frames,
                                     void *pcm_host);
/* Same into device memory, asynchronous. */
FIR_GPU_API int fir_gpu_encode_dev(fir_gpu_ctx *ctx, double scale, void *pcm_dev);

/* ---- the whole path in one call ------------------------------------------- */

/* process_file()'s arithmetic (ProcessFile.cp:41-101,117) for one payload in host
 * memory: apply, peak, the scale rule of ProcessFile.cp:98 (normalise when the peak
 * exceeds 1.0 or `normalize` is set), encode into pcm_out_host.  Same result as
 * fir_gpu_apply + fir_gpu_peak + fir_gpu_encode, bit for bit; what it adds is overlap:
 * without `normalize` it bets that the peak stays <= 1 (no rescaling), encodes each
 * chunk as soon as it is filtered and downloads it under the FIR of the next chunk;
 * only if the final peak exceeds 1 is everything encoded again with the real scale.
 * Synchronous on return; the parked signal stays available. */
FIR_GPU_API int fir_gpu_process(fir_gpu_ctx *ctx, const fir_gpu_kernel *k, const void *pcm_in_host,
                                const fir_gpu_pcm *fmt, int normalize, void *pcm_out_host, double *peak,
                                double *scale);

/* ---- per-phase device times ------------------------------------------------ */

FIR_GPU_API int fir_gpu_last_timing(fir_gpu_ctx *ctx, fir_gpu_timing *t);

/* Measurement, synthetic-input and tuning entry points (probes, kernel variants,
 * caller-supplied taps) are NOT part of the drop-in boundary: they are declared in
 * fir_gpu_dev.h and used only by bench.py, tools/ and the tests. */

#ifdef __cplusplus
}
#endif
#endif /* FIR_GPU_H */
