// fir_gpu.cu -- host side of the C-ABI declared in include/fir_gpu.h.
//
// Device-memory plan per context (one context per GPU):
//   d_pcm   interleaved PCM staging (host entry points only; encode reuses it)
//   d_x     zero-padded planar FP64 input of ONE chunk: [channels][x_pitch],
//           x_pitch = roundup(chunk_frames + 2H, 16); sample j of the chunk sits
//           at index j + H.  Bounded (X_BUDGET) so that hour-long multichannel
//           files stream through it chunk by chunk.
//   d_y     the PARKED filtered signal, planar FP64 [channels][y_pitch]: it has
//           to outlive the FIR because the encode gain depends on the global
//           peak (ProcessFile.cp:92-101), which in sample-block mode is only
//           known after the cross-GPU max.
//   d_peak  8 bytes: bit pattern of max|y| (fused into the FIR epilogue)
// There is no CPU fallback anywhere in this file: every entry point needs a
// live sm_100 device.
#include "../../include/fir_gpu.h"
#include "../../include/fir_gpu_dev.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <numeric>
#include <string>
#include <vector>

#include "fir_dmma.cuh"
#include "fir_fp64.cuh"
#include "pcm_codec.cuh"
#include "sinc_kernel.cuh"

using namespace firgpu;

namespace {

thread_local std::string g_err;

// Test hook (fir_gpu_dev.h): the next n fir_gpu_create calls fail after their streams, events
// and device buffers exist, to prove that a half-built context is torn down completely.
std::atomic<int> fir_gpu_test_fail_create{0};

int fail(int code, const std::string& msg)
{
	g_err = msg;
	return code;
}

#define CU_TRY(expr)                                                                              \
	do {                                                                                          \
		cudaError_t e_ = (expr);                                                                  \
		if (e_ != cudaSuccess) {                                                                  \
			cudaGetLastError(); /* reported here: do not leave it for a later, unrelated call */ \
			return fail(e_ == cudaErrorMemoryAllocation ? FIR_GPU_ERR_NOMEM : FIR_GPU_ERR_CUDA,   \
			            std::string(#expr) + ": " + cudaGetErrorString(e_));                      \
		}                                                                                         \
	} while (0)

inline int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

// ---- FIR kernel variants ------------------------------------------------------

typedef void (*DfmaFn)(const CUtensorMap, const double*, int, double*, long long, long long, unsigned long long*);
typedef void (*DmmaFn)(const double*, long long, const double*, int, int, double*, long long, long long,
                       unsigned long long*);

struct FirVariant {
	const char* name;
	int dmma; // 0: DFMA kernel (fir_fp64.cuh), 1: DMMA kernel (fir_dmma.cuh)
	int nt, kt, t_out, box_rows, smem, ctas_per_sm;
	DfmaFn dfma_kernel;
	DmmaFn dmma_kernel;
};

template <class Cfg>
FirVariant dfma_variant(const char* name)
{
	return {name, 0, Cfg::NT, Cfg::KT, Cfg::T_OUT, Cfg::BOX_ROWS, Cfg::SMEM_BYTES, Cfg::MINB, fir_fp64_kernel<Cfg>, nullptr};
}

template <class Cfg>
FirVariant dmma_variant(const char* name)
{
	return {name, 1, Cfg::NT, Cfg::KT, Cfg::T_OUT, 0, Cfg::SMEM_BYTES, Cfg::MINB, nullptr, fir_dmma_kernel<Cfg>};
}

constexpr int MAX_KT = 1024; // tap arrays are zero-padded generously beyond any variant's tile
constexpr int TAP_PAD = 8;   // zeros in front of h[0] (the DMMA Toeplitz blocks reach k = -7)

const FirVariant* fir_variants(int* n)
{
	// Index 0 is the default: the best across configs 1, 2, 4, 5 in the sweeps on B200
	// (tools/sweep_variants.py, profiles/r1_variant_sweep.txt): 36.9 TFLOP/s on config 2, 36.2 on config 1.
	// The other shapes tried are in the sweep record; these are kept selectable.
	// The product build carries the default and ONE FMA-pipe comparison kernel; the other shapes
	// (2.3 MB of cubin the CLI would page in at start-up for nothing) need -DFIR_ALL_VARIANTS
	// (`make -C audio_fir_filter_b200/csrc sweep`, used by tools/sweep_variants.py).
	static const FirVariant v[] = {
		dmma_variant<DmmaCfg<128, 2, 256, 3, 6>>("dmma_nt128_t2_kt256_s3_b6"),
		dfma_variant<FirCfg<256, 512, 2, 2>>("dfma_nt256_kt512_s2_b2"),
#ifdef FIR_ALL_VARIANTS
		dfma_variant<FirCfg<256, 512, 3, 1>>("dfma_nt256_kt512_s3_b1"),
		dfma_variant<FirCfg<128, 512, 2, 4>>("dfma_nt128_kt512_s2_b4"),
		dmma_variant<DmmaCfg<128, 2, 128, 4, 6>>("dmma_nt128_t2_kt128_s4_b6"),
		dmma_variant<DmmaCfg<128, 2, 64, 6, 6>>("dmma_nt128_t2_kt64_s6_b6"),
		dmma_variant<DmmaCfg<256, 2, 512, 2, 3>>("dmma_nt256_t2_kt512_s2_b3"),
		dmma_variant<DmmaCfg<256, 3, 512, 2, 2>>("dmma_nt256_t3_kt512_s2_b2"),
		dmma_variant<DmmaCfg<128, 3, 128, 4, 4>>("dmma_nt128_t3_kt128_s4_b4"),
		dmma_variant<DmmaCfg<64, 2, 128, 4, 12>>("dmma_nt64_t2_kt128_s4_b12"),
#endif
	};
	*n = (int) (sizeof(v) / sizeof(v[0]));
	return v;
}

const FirVariant& variant_of(const fir_gpu_ctx* c);

// Doubles per channel the FIR of `frames` outputs may read from the padded input.
int64_t x_pitch_for(const FirVariant& v, int64_t frames, int64_t n_taps)
{
	const int64_t H = (n_taps - 1) / 2;
	if (!v.dmma) return (std::max<int64_t>(frames, 1) + 2 * H + 15) / 16 * 16;
	const int64_t blocks = (std::max<int64_t>(frames, 1) + v.t_out - 1) / v.t_out;
	const int64_t n_ktiles = (n_taps + 7 + v.kt - 1) / v.kt;
	return blocks * v.t_out + n_ktiles * v.kt;
}

struct EventPair {
	cudaEvent_t a = nullptr, b = nullptr;
};

} // namespace

// one pending progress callback (cudaLaunchHostFunc payload)
struct ProgressNote {
	fir_gpu_progress_fn fn;
	void* user;
	int64_t done, total;
};

struct fir_gpu_kernel {
	int device = 0;
	int64_t n_taps = 0;   // M + 1
	int64_t n_alloc = 0;  // doubles allocated: TAP_PAD zeros, the taps, a zero tail
	double* d_tpad = nullptr;
	double* d_taps = nullptr; // = d_tpad + TAP_PAD
};

struct fir_gpu_ctx {
	int device = 0;
	int sm_count = 0;
	cudaStream_t own_stream = nullptr, stream = nullptr;
	cudaStream_t copy_stream = nullptr; // H2D of the next chunk runs under the FIR of the current one
	cudaEvent_t copy_done = nullptr, pcm_free = nullptr, feed_last = nullptr;
	// Highest-priority stream for everything on the way OUT (encode to host + its download, the
	// speculative downloads of fir_gpu_process): with several contexts on one GPU the output of
	// one file gets SM slots and the copy engine ahead of the FIR CTAs of the next file.
	cudaStream_t d2h_stream = nullptr;
	cudaEvent_t spec_done = nullptr, enc_ready = nullptr;
	unsigned char* d_out = nullptr;    // encoded PCM of the chunks already filtered (fir_gpu_process)
	size_t out_cap = 0;
	PFN_encodeTiled encode_tiled = nullptr;

	unsigned char* d_pcm = nullptr;
	size_t pcm_cap = 0;
	double* d_x = nullptr;
	size_t x_cap = 0;
	double* d_y = nullptr;
	size_t y_cap = 0;
	unsigned long long* d_peak = nullptr;
	double* d_sink = nullptr;
	dd* d_lp = nullptr; // build_kernel scratch: low-pass taps (double-double) + block partial sums
	size_t lp_cap = 0;

	bool parked = false;

	// the apply in progress (fir_gpu_apply_begin .. _end; fir_gpu_apply is begin + feed + end)
	struct Pass {
		bool active = false, from_host = false;
		const fir_gpu_kernel* k = nullptr;
		const unsigned char* pcm_dev = nullptr; // interleaved PCM on the device, byte 0 = frame -halo_left
		std::vector<std::pair<int64_t, int64_t>> chunks; // (first output frame, frames)
		size_t next = 0;       // next chunk to launch
		size_t bytes_fed = 0;  // of the host payload, uploaded or in flight
		size_t bytes_total = 0;
		int64_t done_frames = 0;
		unsigned char* spec_out = nullptr; // host buffer that receives each chunk encoded with scale 1 (or null)
	} pass;
	fir_gpu_progress_fn progress = nullptr;
	void* progress_user = nullptr;
	fir_gpu_pcm fmt{};
	int64_t y_pitch = 0;
	int variant = 0;

	// timing of the last apply / encode
	std::vector<EventPair> pool;
	size_t pool_used = 0;
	std::vector<size_t> t_h2d, t_decode, t_fir, t_peak, t_encode, t_d2h;
	int64_t fir_launches = 0, other_launches = 0;
	int64_t x_budget_bytes = (int64_t) 2 << 30;
	// cudaFuncSetAttribute is per device: remember what this context's device has
	struct CodecSlot {
		bool attr_done = false;
		int nt = 0, resident = 0, carveout = -2;
		unsigned smem = 0;
	} codec_slot[12];
	int codec_carveout = CODEC_CARVEOUT; // percent of the SM's array given to shared memory (-1: driver's choice)
	int codec_tile_bytes = CODEC_TILE_BYTES, codec_nt = CODEC_NT;
	std::vector<char> fir_attr;
};

namespace {

const FirVariant& variant_of(const fir_gpu_ctx* c)
{
	int nv = 0;
	const FirVariant* vs = fir_variants(&nv);
	return vs[(c->variant >= 0 && c->variant < nv) ? c->variant : 0];
}

struct DeviceGuard {
	int prev = -1;
	explicit DeviceGuard(int dev)
	{
		cudaGetDevice(&prev);
		if (prev != dev) cudaSetDevice(dev);
		else prev = -1;
	}
	~DeviceGuard()
	{
		if (prev >= 0) cudaSetDevice(prev);
	}
};

int ensure(void** p, size_t* cap, size_t need)
{
	if (*cap >= need) return FIR_GPU_OK;
	if (*p) cudaFree(*p);
	*p = nullptr;
	*cap = 0;
	// grow with a little slack; +64 so codec tails can be staged safely
	const size_t want = need + need / 16 + 256;
	cudaError_t e = cudaMalloc(p, want);
	if (e != cudaSuccess) {
		cudaGetLastError();
		e = cudaMalloc(p, need + 256);
		if (e != cudaSuccess) {
			cudaGetLastError(); // clear it: the failure is reported through the status code
			*p = nullptr;
			return fail(FIR_GPU_ERR_NOMEM, std::string("cudaMalloc(") + std::to_string(need) +
			                                   " B): " + cudaGetErrorString(e));
		}
		*cap = need + 256;
		return FIR_GPU_OK;
	}
	*cap = want;
	return FIR_GPU_OK;
}

// A timing span = two events of the context's pool around a piece of work.  Creating or
// recording an event can fail (out of resources, a sticky error on the context): reported,
// not ignored.
int begin_span(fir_gpu_ctx* c, size_t* idx, cudaStream_t st = nullptr)
{
	if (c->pool_used == c->pool.size()) {
		EventPair p;
		CU_TRY(cudaEventCreate(&p.a));
		cudaError_t e = cudaEventCreate(&p.b);
		if (e != cudaSuccess) {
			cudaGetLastError();
			cudaEventDestroy(p.a);
			return fail(FIR_GPU_ERR_CUDA, std::string("cudaEventCreate: ") + cudaGetErrorString(e));
		}
		c->pool.push_back(p);
	}
	CU_TRY(cudaEventRecord(c->pool[c->pool_used].a, st ? st : c->stream));
	*idx = c->pool_used++;
	return FIR_GPU_OK;
}

int end_span(fir_gpu_ctx* c, size_t i, cudaStream_t st = nullptr)
{
	CU_TRY(cudaEventRecord(c->pool[i].b, st ? st : c->stream));
	return FIR_GPU_OK;
}

#define SPAN_BEGIN(var, ...)                                  \
	size_t var = 0;                                           \
	do {                                                      \
		int rc_ = begin_span(c, &var, ##__VA_ARGS__);         \
		if (rc_) return rc_;                                  \
	} while (0)
#define SPAN_END(var, ...)                                    \
	do {                                                      \
		int rc_ = end_span(c, var, ##__VA_ARGS__);            \
		if (rc_) return rc_;                                  \
	} while (0)

void reset_timing(fir_gpu_ctx* c, bool all)
{
	if (all) {
		c->pool_used = 0;
		c->t_h2d.clear();
		c->t_decode.clear();
		c->t_fir.clear();
		c->t_peak.clear();
		c->fir_launches = 0;
		c->other_launches = 0;
		c->t_encode.clear();
		c->t_d2h.clear();
	}
}

double span_ms(fir_gpu_ctx* c, const std::vector<size_t>& v)
{
	double tot = 0.0;
	for (size_t i : v) {
		float ms = 0.f;
		if (cudaEventElapsedTime(&ms, c->pool[i].a, c->pool[i].b) == cudaSuccess) tot += ms;
	}
	return tot;
}

int check_fmt(const fir_gpu_pcm* f)
{
	if (!f) return fail(FIR_GPU_ERR_INVALID, "null format");
	if (f->frames < 0 || f->halo_left < 0 || f->halo_right < 0)
		return fail(FIR_GPU_ERR_INVALID, "negative frame count");
	if (f->channels < 1 || f->channels > 256) return fail(FIR_GPU_ERR_INVALID, "channels must be 1..256");
	if (f->bits != 16 && f->bits != 24 && f->bits != 32)
		return fail(FIR_GPU_ERR_INVALID, "bits must be 16, 24 or 32");
	return FIR_GPU_OK;
}

// The tile counter of a codec launch (pcm_codec.cuh: next_tile).  Launches that may overlap on
// the device never share one: decode and encode have their own, and so has each stream.
unsigned long long* tile_counter_for(fir_gpu_ctx* c, int encode, cudaStream_t st)
{
	const int stream_slot = st == c->d2h_stream ? 1 : 0;
	return c->d_peak + 2 + 2 * encode + stream_slot; // slots 2..5 of the 64-byte scratch (0: peak, 1: peak recompute)
}

// CTAs of a codec kernel that are resident per SM at this geometry -- asked of the runtime, not
// estimated: the persistent grid must be exactly one wave (a CTA that does not fit starts when
// another finishes, i.e. at the very end, and runs its share of the tiles almost alone).
int codec_residency(fir_gpu_ctx* c, int slot, const void* fn, const CodecGeom& g, int* resident)
{
	fir_gpu_ctx::CodecSlot& s = c->codec_slot[slot];
	if (!s.attr_done) {
		CU_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, CODEC_SMEM_MAX));
		s.attr_done = true;
	}
	if (s.carveout != c->codec_carveout) {
		CU_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, c->codec_carveout));
		s.carveout = c->codec_carveout;
		s.nt = 0; // residency depends on it
	}
	if (s.nt != g.nt || s.smem != g.smem) {
		int n = 0;
		CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, g.nt, g.smem));
		s.nt = g.nt;
		s.smem = g.smem;
		s.resident = n < 1 ? 1 : n;
	}
	*resident = s.resident;
	return FIR_GPU_OK;
}

template <int BITS, bool BE>
int launch_decode(fir_gpu_ctx* c, cudaStream_t st, const unsigned char* pcm, int64_t avail_lo, int64_t avail_hi,
                  int64_t g0, int64_t n_x, int ch, double* x, int64_t x_pitch)
{
	const int fb = ch * (BITS / 8);
	const CodecGeom g = codec_geom(fb, ch, c->codec_tile_bytes, c->codec_nt);
	int resident = 0;
	int rc = codec_residency(c, (BITS / 8 - 2) * 2 + (BE ? 1 : 0), (const void*) pcm_decode_kernel<BITS, BE>, g, &resident);
	if (rc) return rc;
	const int64_t tiles = (n_x + g.frames - 1) / g.frames;
	const unsigned blocks = (unsigned) std::min<int64_t>(tiles, (int64_t) c->sm_count * resident);
	pcm_decode_kernel<BITS, BE><<<blocks, g.nt, g.smem, st>>>(pcm, avail_lo, avail_hi, g0, n_x, ch, x, x_pitch,
	                                                          g.frames, tile_counter_for(c, 0, st));
	CU_TRY(cudaGetLastError());
	c->other_launches++;
	return FIR_GPU_OK;
}

template <int BITS, bool BE>
int launch_encode(fir_gpu_ctx* c, cudaStream_t st, const double* y, int64_t y_pitch, int64_t frames, int ch,
                  double gain, unsigned char* pcm)
{
	const int fb = ch * (BITS / 8);
	const CodecGeom g = codec_geom(fb, ch, c->codec_tile_bytes, c->codec_nt);
	int resident = 0;
	int rc = codec_residency(c, 6 + (BITS / 8 - 2) * 2 + (BE ? 1 : 0), (const void*) pcm_encode_kernel<BITS, BE>, g, &resident);
	if (rc) return rc;
	const int64_t tiles = (frames + g.frames - 1) / g.frames;
	const unsigned blocks = (unsigned) std::min<int64_t>(tiles, (int64_t) c->sm_count * resident);
	pcm_encode_kernel<BITS, BE><<<blocks, g.nt, g.smem, st>>>(y, y_pitch, frames, ch, gain, pcm, g.frames,
	                                                          tile_counter_for(c, 1, st));
	CU_TRY(cudaGetLastError());
	c->other_launches++;
	return FIR_GPU_OK;
}

// rc = fn<BITS, BE>(args...) for the format at hand
#define DISPATCH_CODEC(rc, fn, bits, be, ...)                                  \
	do {                                                                       \
		if ((bits) == 16) rc = (be) ? fn<16, true>(__VA_ARGS__) : fn<16, false>(__VA_ARGS__);      \
		else if ((bits) == 24) rc = (be) ? fn<24, true>(__VA_ARGS__) : fn<24, false>(__VA_ARGS__); \
		else rc = (be) ? fn<32, true>(__VA_ARGS__) : fn<32, false>(__VA_ARGS__);                   \
	} while (0)

// One FIR launch over a zero-padded planar chunk resident at d_x (pitch from x_pitch_for).
int launch_fir(fir_gpu_ctx* c, const fir_gpu_kernel* k, const double* d_x, int64_t x_pitch, int ch, double* y,
               int64_t y_pitch, int64_t frames, unsigned long long* peak)
{
	int nv = 0;
	const FirVariant* vs = fir_variants(&nv);
	const FirVariant& v = variant_of(c);
	const int vi = (int) (&v - vs);
	if ((int) c->fir_attr.size() < nv) c->fir_attr.assign(nv, 0);
	if (!c->fir_attr[vi]) {
		const void* fn = v.dmma ? (const void*) v.dmma_kernel : (const void*) v.dfma_kernel;
		CU_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, v.smem));
		c->fir_attr[vi] = 1;
	}
	dim3 grid((unsigned) ((frames + v.t_out - 1) / v.t_out), (unsigned) ch);
	if (v.dmma) {
		const int n_ktiles = (int) ((k->n_taps + 7 + v.kt - 1) / v.kt);
		const int n_steps = (int) ((k->n_taps - 1 + 7) / 8 + 1); // 8-tap steps that touch a real tap
		v.dmma_kernel<<<grid, v.nt, v.smem, c->stream>>>(d_x, (long long) x_pitch, k->d_tpad, n_ktiles, n_steps, y,
		                                               (long long) y_pitch, (long long) frames, peak);
	} else {
		CUtensorMap map;
		const cuuint64_t dims[3] = {16, (cuuint64_t) (x_pitch / 16), (cuuint64_t) ch};
		const cuuint64_t strides[2] = {128, (cuuint64_t) x_pitch * 8};
		const cuuint32_t box[3] = {16, (cuuint32_t) v.box_rows, 1};
		const cuuint32_t estr[3] = {1, 1, 1};
		CUresult r = c->encode_tiled(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*) d_x, dims, strides, box, estr,
		                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
		                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
		if (r != CUDA_SUCCESS)
			return fail(FIR_GPU_ERR_CUDA, "cuTensorMapEncodeTiled failed: " + std::to_string((int) r));
		const int n_ktiles = (int) ((k->n_taps + v.kt - 1) / v.kt);
		v.dfma_kernel<<<grid, v.nt, v.smem, c->stream>>>(map, k->d_taps, n_ktiles, y, (long long) y_pitch,
		                                               (long long) frames, peak);
	}
	CU_TRY(cudaGetLastError());
	c->fir_launches++;
	return FIR_GPU_OK;
}

int usable_device(int dev, cudaDeviceProp* prop)
{
	cudaDeviceProp p;
	if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	if (prop) *prop = p;
	return p.major == 10; // sm_100 family only: the cubin is sm_100a
}

} // namespace

// ------------------------------------------------------------------ life cycle

extern "C" {

int fir_gpu_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	int ok = 0;
	for (int d = 0; d < n; ++d) ok += usable_device(d, nullptr);
	return ok;
}

const char* fir_gpu_last_error(void) { return g_err.c_str(); }

int fir_gpu_create(int device, fir_gpu_ctx** out)
{
	if (!out) return fail(FIR_GPU_ERR_INVALID, "null out");
	*out = nullptr;
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n == 0) {
		cudaGetLastError();
		return fail(FIR_GPU_ERR_NO_DEVICE,
		            std::string("no CUDA device (") + cudaGetErrorString(e) + "); this library has no CPU path");
	}
	if (device < 0 || device >= n) return fail(FIR_GPU_ERR_INVALID, "device index out of range");
	cudaDeviceProp prop;
	if (!usable_device(device, &prop))
		return fail(FIR_GPU_ERR_NO_DEVICE, "device is not sm_100 (B200); kernels are built for sm_100a only");

	DeviceGuard g(device);
	// a context that fails half-way is torn down again: nothing leaks
	struct Destroyer {
		void operator()(fir_gpu_ctx* p) const { fir_gpu_destroy(p); }
	};
	std::unique_ptr<fir_gpu_ctx, Destroyer> guard(new fir_gpu_ctx());
	fir_gpu_ctx* c = guard.get();
	c->device = device;
	c->sm_count = prop.multiProcessorCount;
	CU_TRY(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
	c->stream = c->own_stream;
	CU_TRY(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
	CU_TRY(cudaEventCreateWithFlags(&c->copy_done, cudaEventDisableTiming));
	CU_TRY(cudaEventCreateWithFlags(&c->pcm_free, cudaEventDisableTiming));
	CU_TRY(cudaEventCreateWithFlags(&c->feed_last, cudaEventDisableTiming));
	int prio_lo = 0, prio_hi = 0;
	CU_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
	CU_TRY(cudaStreamCreateWithPriority(&c->d2h_stream, cudaStreamNonBlocking, prio_hi));
	CU_TRY(cudaEventCreateWithFlags(&c->spec_done, cudaEventDisableTiming));
	CU_TRY(cudaEventCreateWithFlags(&c->enc_ready, cudaEventDisableTiming));
	CU_TRY(cudaMalloc(&c->d_peak, 64));
	CU_TRY(cudaMemset(c->d_peak, 0, 64));
	CU_TRY(cudaMalloc(&c->d_sink, 64));
	if (fir_gpu_test_fail_create.load() > 0 && fir_gpu_test_fail_create.fetch_sub(1) > 0)
		return fail(FIR_GPU_ERR_CUDA, "fir_gpu_create: injected failure (test hook)");

	void* fn = nullptr;
	cudaDriverEntryPointQueryResult qr;
	e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
	if (e != cudaSuccess || qr != cudaDriverEntryPointSuccess || !fn)
		return fail(FIR_GPU_ERR_NO_DEVICE, "driver lacks cuTensorMapEncodeTiled (TMA)");
	c->encode_tiled = (PFN_encodeTiled) fn;
	*out = guard.release();
	return FIR_GPU_OK;
}

void fir_gpu_destroy(fir_gpu_ctx* c)
{
	if (!c) return;
	DeviceGuard g(c->device);
	if (c->stream) cudaStreamSynchronize(c->stream);
	if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
	if (c->d2h_stream) cudaStreamSynchronize(c->d2h_stream);
	for (auto& p : c->pool) {
		cudaEventDestroy(p.a);
		cudaEventDestroy(p.b);
	}
	cudaFree(c->d_pcm);
	cudaFree(c->d_x);
	cudaFree(c->d_y);
	cudaFree(c->d_peak);
	cudaFree(c->d_sink);
	cudaFree(c->d_lp);
	if (c->own_stream) cudaStreamDestroy(c->own_stream);
	if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
	if (c->copy_done) cudaEventDestroy(c->copy_done);
	if (c->pcm_free) cudaEventDestroy(c->pcm_free);
	if (c->feed_last) cudaEventDestroy(c->feed_last);
	if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
	if (c->spec_done) cudaEventDestroy(c->spec_done);
	if (c->enc_ready) cudaEventDestroy(c->enc_ready);
	cudaFree(c->d_out);
	cudaGetLastError(); // teardown never leaves an error behind for an unrelated call
	delete c;
}

int fir_gpu_set_stream(fir_gpu_ctx* c, void* s)
{
	if (!c) return fail(FIR_GPU_ERR_INVALID, "null context");
	c->stream = s ? (cudaStream_t) s : c->own_stream;
	return FIR_GPU_OK;
}

int fir_gpu_synchronize(fir_gpu_ctx* c)
{
	if (!c) return fail(FIR_GPU_ERR_INVALID, "null context");
	DeviceGuard g(c->device);
	CU_TRY(cudaStreamSynchronize(c->stream));
	return FIR_GPU_OK;
}

void* fir_gpu_host_alloc(size_t bytes)
{
	void* p = nullptr;
	if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
		cudaGetLastError();
		g_err = "cudaHostAlloc failed";
		return nullptr;
	}
	return p;
}

void fir_gpu_host_free(void* p)
{
	if (p) cudaFreeHost(p);
}

int fir_gpu_set_variant(fir_gpu_ctx* c, int variant)
{
	if (!c) return fail(FIR_GPU_ERR_INVALID, "null context");
	int nv = 0;
	fir_variants(&nv);
	if (variant < 0 || variant >= nv) return fail(FIR_GPU_ERR_INVALID, "no such FIR variant");
	c->variant = variant;
	return FIR_GPU_OK;
}

int fir_gpu_variant_count(void)
{
	int nv = 0;
	fir_variants(&nv);
	return nv;
}

const char* fir_gpu_variant_name(int variant)
{
	int nv = 0;
	const FirVariant* vs = fir_variants(&nv);
	return (variant >= 0 && variant < nv) ? vs[variant].name : "";
}

int fir_gpu_set_x_budget(fir_gpu_ctx* c, int64_t bytes)
{
	if (!c) return fail(FIR_GPU_ERR_INVALID, "null context");
	if (bytes < (1 << 20)) return fail(FIR_GPU_ERR_INVALID, "input scratch budget must be at least 1 MiB");
	c->x_budget_bytes = bytes;
	return FIR_GPU_OK;
}

// ------------------------------------------------------------ build_kernel

static int alloc_kernel(fir_gpu_ctx* c, int64_t n_taps, fir_gpu_kernel** out)
{
	fir_gpu_kernel* k = new fir_gpu_kernel();
	k->device = c->device;
	k->n_taps = n_taps;
	k->n_alloc = TAP_PAD + round_up(n_taps + 8, MAX_KT) + MAX_KT + 32;
	cudaError_t e = cudaMalloc(&k->d_tpad, (size_t) k->n_alloc * sizeof(double));
	if (e == cudaSuccess) e = cudaMemsetAsync(k->d_tpad, 0, (size_t) k->n_alloc * sizeof(double), c->stream);
	if (e != cudaSuccess) {
		cudaGetLastError();
		cudaFree(k->d_tpad);
		delete k;
		return fail(FIR_GPU_ERR_NOMEM, std::string("cudaMalloc taps: ") + cudaGetErrorString(e));
	}
	k->d_taps = k->d_tpad + TAP_PAD;
	*out = k;
	return FIR_GPU_OK;
}

int fir_gpu_build_kernel(fir_gpu_ctx* c, double fc_norm, double bw_norm, fir_gpu_kernel** out, int64_t* half_len)
{
	if (!c || !out) return fail(FIR_GPU_ERR_INVALID, "null argument");
	*out = nullptr;
	if (!(bw_norm > 0.0) || !std::isfinite(bw_norm)) return fail(FIR_GPU_ERR_INVALID, "slope must be > 0");
	if (!(fc_norm > 0.0) || !(fc_norm < 0.5))
		return fail(FIR_GPU_ERR_INVALID, "cutoff must lie in (0, sampleRate/2)");
	// M = round(4/bw), forced even (decision D4)
	const long double mf = 4.0L / (long double) bw_norm;
	if (mf > 1.0e9L) return fail(FIR_GPU_ERR_INVALID, "slope too narrow: more than 1e9 taps");
	int64_t M = (int64_t) llroundl(mf);
	if (M & 1) ++M;
	if (M < 2) M = 2;

	DeviceGuard g(c->device);
	fir_gpu_kernel* k = nullptr;
	int rc = alloc_kernel(c, M + 1, &k);
	if (rc) return rc;
	const int blocks = (int) ((M + 1 + 255) / 256);
	// scratch kept by the context: a cudaFree per call would synchronise the whole device
	// (and every other context working on it) once per file
	rc = ensure((void**) &c->d_lp, &c->lp_cap, (size_t) (M + 1 + blocks + 1) * sizeof(dd));
	if (rc) {
		fir_gpu_kernel_free(k);
		return rc;
	}
	dd* d_lp = c->d_lp;
	dd* d_part = c->d_lp + (M + 1);
	sinc_lowpass_kernel<<<blocks, 256, 0, c->stream>>>(M, fc_norm, d_lp, d_part);
	sinc_sum_kernel<<<1, 256, 0, c->stream>>>(d_part, blocks, d_part + blocks);
	sinc_lowcut_kernel<<<blocks, 256, 0, c->stream>>>(M, d_lp, d_part + blocks, k->d_taps, M + 1);
	cudaError_t e = cudaGetLastError();
	if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
	if (e != cudaSuccess) {
		fir_gpu_kernel_free(k);
		return fail(FIR_GPU_ERR_CUDA, std::string("build_kernel: ") + cudaGetErrorString(e));
	}
	c->other_launches += 3;
	if (half_len) *half_len = M / 2;
	*out = k;
	return FIR_GPU_OK;
}

int fir_gpu_kernel_from_taps(fir_gpu_ctx* c, const double* taps, int64_t n_taps, fir_gpu_kernel** out)
{
	if (!c || !out || !taps) return fail(FIR_GPU_ERR_INVALID, "null argument");
	*out = nullptr;
	if (n_taps < 1 || !(n_taps & 1)) return fail(FIR_GPU_ERR_INVALID, "tap count must be odd (M even)");
	DeviceGuard g(c->device);
	fir_gpu_kernel* k = nullptr;
	int rc = alloc_kernel(c, n_taps, &k);
	if (rc) return rc;
	cudaError_t e = cudaMemcpyAsync(k->d_taps, taps, (size_t) n_taps * sizeof(double), cudaMemcpyHostToDevice,
	                                c->stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
	if (e != cudaSuccess) {
		fir_gpu_kernel_free(k);
		return fail(FIR_GPU_ERR_CUDA, std::string("kernel_from_taps: ") + cudaGetErrorString(e));
	}
	*out = k;
	return FIR_GPU_OK;
}

int64_t fir_gpu_kernel_num_taps(const fir_gpu_kernel* k) { return k ? k->n_taps : 0; }

int fir_gpu_kernel_taps(fir_gpu_ctx* c, const fir_gpu_kernel* k, double* taps_out, int64_t n)
{
	if (!c || !k || !taps_out) return fail(FIR_GPU_ERR_INVALID, "null argument");
	if (n != k->n_taps) return fail(FIR_GPU_ERR_INVALID, "tap count mismatch");
	DeviceGuard g(c->device);
	CU_TRY(cudaMemcpyAsync(taps_out, k->d_taps, (size_t) n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
	CU_TRY(cudaStreamSynchronize(c->stream));
	return FIR_GPU_OK;
}

void fir_gpu_kernel_free(fir_gpu_kernel* k)
{
	if (!k) return;
	DeviceGuard g(k->device);
	cudaFree(k->d_tpad);
	delete k;
}

// ------------------------------------------------------------------- apply

// The chunks of one apply: [f0, f0+nf) output frames each.  Every chunk fits the
// decoded-input budget.  From host memory the first chunk is a short one, so that the
// upload of everything else hides under its FIR (and under the FIRs that follow);
// a streamed apply (feed by feed) is cut into ~16 chunks that start as their bytes land.
// (Measured in round 2: cutting the host path into 8 whole-wave chunks instead costs 0.85 ms of a
// 53 ms config-2 pass -- every kernel boundary drains and refills the SMs -- and only pays when
// eight ranks share one host's PCIe; hosts that want the copies hidden keep two files in flight.)
static std::vector<std::pair<int64_t, int64_t>> plan_chunks(const fir_gpu_ctx* c, const FirVariant& v, int64_t frames,
                                                            int ch, int64_t n_taps,
                                                            int mode /*0 dev, 1 host, 2 stream, 3 host + speculative out*/)
{
	const int64_t t_out = v.t_out;
	int64_t chunk = (c->x_budget_bytes / 8 / ch - (n_taps + 2 * MAX_KT)) / t_out * t_out;
	if (chunk < t_out) chunk = t_out;
	std::vector<std::pair<int64_t, int64_t>> out;
	int64_t f0 = 0;
	if (mode == 2 && frames >= 256 * t_out) chunk = std::min(chunk, round_up(frames / 16, t_out));
	if (mode == 3) {
		// speculative downloads: one chunk per ~2.5 ms of FIR (35 TFLOP/s), at most 8, each a whole
		// number of full waves of CTAs so that the extra launches do not add partial-wave tails
		const double est_ms = 2.0 * (double) n_taps * (double) frames * ch / 35.0e9;
		const int64_t n = std::clamp<int64_t>((int64_t) (est_ms / 2.5), 1, 8);
		const int64_t wave = (int64_t) c->sm_count * v.ctas_per_sm;
		const int64_t bx_unit = std::max<int64_t>(1, wave / std::gcd<int64_t>(wave, ch)); // grid.x per whole waves
		const int64_t bx = round_up((frames / n + t_out - 1) / t_out, bx_unit);
		if (n > 1) chunk = std::min(chunk, bx * t_out);
		else mode = 1; // too short to split for the downloads: at least hide the upload
	}
	if (mode == 1 && frames >= 256 * t_out) {
		const int64_t first = std::min(chunk, round_up(frames / 16, t_out));
		out.emplace_back(0, first);
		f0 = first;
	}
	for (; f0 < frames; f0 += chunk) out.emplace_back(f0, std::min(chunk, frames - f0));
	return out;
}

static void CUDART_CB progress_trampoline(void* p)
{
	// the callback owns its note: nothing else has to guess when the stream got this far
	std::unique_ptr<ProgressNote> n(static_cast<ProgressNote*>(p));
	n->fn(n->done, n->total, n->user); // runs on a CUDA callback thread: no CUDA calls in there
}

static int pass_begin(fir_gpu_ctx* c, const fir_gpu_kernel* k, const unsigned char* pcm_dev, const fir_gpu_pcm* fmt,
                      int mode)
{
	if (k->device != c->device) return fail(FIR_GPU_ERR_STATE, "kernel lives on another device");
	const int ch = fmt->channels;
	const int64_t frames = fmt->frames;
	c->parked = false;
	c->fmt = *fmt;
	c->y_pitch = round_up(std::max<int64_t>(frames, 1), 16);
	int rc = ensure((void**) &c->d_y, &c->y_cap, (size_t) c->y_pitch * ch * sizeof(double));
	if (rc) return rc;
	CU_TRY(cudaMemsetAsync(c->d_peak, 0, 8, c->stream));

	fir_gpu_ctx::Pass& p = c->pass;
	p = fir_gpu_ctx::Pass();
	p.active = true;
	p.from_host = mode != 0;
	p.k = k;
	p.pcm_dev = pcm_dev;
	p.chunks = plan_chunks(c, variant_of(c), frames, ch, k->n_taps, mode);
	// size the decoded-input scratch for the largest chunk now: growing it between chunks
	// would mean a cudaFree, i.e. a device-wide synchronisation in the middle of the pass
	int64_t max_pitch = 0;
	for (const auto& [f0, nf] : p.chunks) max_pitch = std::max(max_pitch, x_pitch_for(variant_of(c), nf, k->n_taps));
	rc = ensure((void**) &c->d_x, &c->x_cap, (size_t) max_pitch * ch * sizeof(double));
	if (rc) return rc;
	p.bytes_total = (size_t) (fmt->halo_left + frames + fmt->halo_right) * ch * (fmt->bits / 8);
	if (p.from_host) {
		// the staging buffer may still be read by earlier work of the compute stream
		CU_TRY(cudaEventRecord(c->pcm_free, c->stream));
		CU_TRY(cudaStreamWaitEvent(c->copy_stream, c->pcm_free, 0));
	}
	return FIR_GPU_OK;
}

// Upload the next n bytes of the payload on the copy stream.
static int pass_upload(fir_gpu_ctx* c, const unsigned char* src, size_t n)
{
	fir_gpu_ctx::Pass& p = c->pass;
	if (!n) return FIR_GPU_OK;
	SPAN_BEGIN(s, c->copy_stream);
	CU_TRY(cudaMemcpyAsync(const_cast<unsigned char*>(p.pcm_dev) + p.bytes_fed, src, n, cudaMemcpyHostToDevice,
	                       c->copy_stream));
	SPAN_END(s, c->copy_stream);
	c->t_h2d.push_back(s);
	p.bytes_fed += n;
	return FIR_GPU_OK;
}

// Launch decode + FIR of every chunk whose samples (incl. halo) have been uploaded.
static int pass_launch_ready(fir_gpu_ctx* c)
{
	fir_gpu_ctx::Pass& p = c->pass;
	const fir_gpu_pcm& fmt = c->fmt;
	const fir_gpu_kernel* k = p.k;
	const FirVariant& v = variant_of(c);
	const int ch = fmt.channels;
	const size_t fb = (size_t) ch * (fmt.bits / 8);
	const int64_t H = (k->n_taps - 1) / 2;
	const int64_t avail_lo = -fmt.halo_left, avail_hi = fmt.frames + fmt.halo_right;
	const int64_t uploaded = p.from_host ? avail_lo + (int64_t) (p.bytes_fed / fb) : avail_hi;
	bool waited = false;
	while (p.next < p.chunks.size()) {
		const auto [f0, nf] = p.chunks[p.next];
		if (std::min(avail_hi, f0 + nf + H) > uploaded) break;
		if (p.from_host && !waited) {
			CU_TRY(cudaEventRecord(c->copy_done, c->copy_stream));
			CU_TRY(cudaStreamWaitEvent(c->stream, c->copy_done, 0));
			waited = true;
		}
		const int64_t x_pitch = x_pitch_for(v, nf, k->n_taps);
		int rc = ensure((void**) &c->d_x, &c->x_cap, (size_t) x_pitch * ch * sizeof(double));
		if (rc) return rc;
		SPAN_BEGIN(sd);
		DISPATCH_CODEC(rc, launch_decode, fmt.bits, fmt.big_endian != 0, c, c->stream, p.pcm_dev, avail_lo, avail_hi,
		               f0 - H, x_pitch, ch, c->d_x, x_pitch);
		if (rc) return rc;
		SPAN_END(sd);
		c->t_decode.push_back(sd);
		SPAN_BEGIN(sf);
		rc = launch_fir(c, k, c->d_x, x_pitch, ch, c->d_y + f0, c->y_pitch, nf, c->d_peak);
		if (rc) return rc;
		SPAN_END(sf);
		c->t_fir.push_back(sf);
		if (p.spec_out) {
			// speculate that the file needs no rescaling (peak <= 1, no -n): encode this chunk
			// with scale 1 as soon as it is filtered (on the high-priority output stream, so it
			// does not queue behind the next chunk's FIR) and download it under that FIR
			unsigned char* dst = c->d_out + (size_t) f0 * fb;
			CU_TRY(cudaEventRecord(c->spec_done, c->stream));
			CU_TRY(cudaStreamWaitEvent(c->d2h_stream, c->spec_done, 0));
			SPAN_BEGIN(se, c->d2h_stream);
			DISPATCH_CODEC(rc, launch_encode, fmt.bits, fmt.big_endian != 0, c, c->d2h_stream, c->d_y + f0, c->y_pitch,
			               nf, ch, std::ldexp(1.0, fmt.bits - 1), dst);
			if (rc) return rc;
			SPAN_END(se, c->d2h_stream);
			c->t_encode.push_back(se);
			SPAN_BEGIN(sc, c->d2h_stream);
			CU_TRY(cudaMemcpyAsync(p.spec_out + (size_t) f0 * fb, dst, (size_t) nf * fb, cudaMemcpyDeviceToHost,
			                       c->d2h_stream));
			SPAN_END(sc, c->d2h_stream);
			c->t_d2h.push_back(sc);
		}
		p.done_frames = f0 + nf;
		if (c->progress) {
			ProgressNote* n = new ProgressNote{c->progress, c->progress_user, p.done_frames, fmt.frames};
			const cudaError_t pe = cudaLaunchHostFunc(c->stream, progress_trampoline, n);
			if (pe != cudaSuccess) {
				delete n; // never enqueued
				CU_TRY(pe);
			}
		}
		++p.next;
	}
	return FIR_GPU_OK;
}

static int pass_end(fir_gpu_ctx* c)
{
	fir_gpu_ctx::Pass& p = c->pass;
	if (p.from_host && p.bytes_fed != p.bytes_total) {
		p.active = false;
		return fail(FIR_GPU_ERR_STATE, "fir_gpu_apply_end: " + std::to_string(p.bytes_fed) + " of " +
		                                   std::to_string(p.bytes_total) + " payload bytes were fed");
	}
	int rc = pass_launch_ready(c);
	p.active = false;
	if (rc) return rc;
	c->parked = true;
	return FIR_GPU_OK;
}

static int check_apply_args(fir_gpu_ctx* c, const fir_gpu_kernel* k, const fir_gpu_pcm* fmt)
{
	if (!c || !k) return fail(FIR_GPU_ERR_INVALID, "null argument");
	return check_fmt(fmt);
}

// Upload the host payload range by range, each just ahead of the chunk that needs it.
static int pass_feed_all(fir_gpu_ctx* c, const unsigned char* src, const fir_gpu_pcm* fmt, size_t in_bytes)
{
	const size_t fb = (size_t) fmt->channels * (fmt->bits / 8);
	const int64_t H = (c->pass.k->n_taps - 1) / 2;
	for (const auto& [f0, nf] : c->pass.chunks) {
		const int64_t need = std::min(fmt->frames + fmt->halo_right, f0 + nf + H) + fmt->halo_left; // frames from byte 0
		const size_t upto = std::min(in_bytes, (size_t) need * fb);
		if (upto > c->pass.bytes_fed) {
			int rc = pass_upload(c, src + c->pass.bytes_fed, upto - c->pass.bytes_fed);
			if (!rc) rc = pass_launch_ready(c);
			if (rc) return rc;
		}
	}
	if (c->pass.bytes_fed < in_bytes) // halo_right beyond half_len: not needed by any chunk
		return pass_upload(c, src + c->pass.bytes_fed, in_bytes - c->pass.bytes_fed);
	return FIR_GPU_OK;
}

int fir_gpu_apply(fir_gpu_ctx* c, const fir_gpu_kernel* k, const void* pcm_host, const fir_gpu_pcm* fmt)
{
	int rc = check_apply_args(c, k, fmt);
	if (rc) return rc;
	if (!pcm_host && fmt->frames > 0) return fail(FIR_GPU_ERR_INVALID, "null PCM buffer");
	DeviceGuard g(c->device);
	reset_timing(c, true);
	const size_t fb = (size_t) fmt->channels * (fmt->bits / 8);
	const size_t in_bytes = (size_t) (fmt->halo_left + fmt->frames + fmt->halo_right) * fb;
	rc = ensure((void**) &c->d_pcm, &c->pcm_cap, in_bytes + 32);
	if (rc) return rc;
	rc = pass_begin(c, k, c->d_pcm, fmt, 1);
	if (rc) return rc;
	rc = pass_feed_all(c, static_cast<const unsigned char*>(pcm_host), fmt, in_bytes);
	if (rc) return rc;
	return pass_end(c);
}

int fir_gpu_apply_dev(fir_gpu_ctx* c, const fir_gpu_kernel* k, const void* pcm_dev, const fir_gpu_pcm* fmt)
{
	int rc = check_apply_args(c, k, fmt);
	if (rc) return rc;
	if (!pcm_dev && fmt->frames > 0) return fail(FIR_GPU_ERR_INVALID, "null PCM buffer");
	DeviceGuard g(c->device);
	reset_timing(c, true);
	rc = pass_begin(c, k, static_cast<const unsigned char*>(pcm_dev), fmt, 0);
	if (rc) return rc;
	return pass_end(c);
}

int fir_gpu_process(fir_gpu_ctx* c, const fir_gpu_kernel* k, const void* pcm_in_host, const fir_gpu_pcm* fmt,
                    int normalize, void* pcm_out_host, double* peak_out, double* scale_out)
{
	int rc = check_apply_args(c, k, fmt);
	if (rc) return rc;
	if ((!pcm_in_host || !pcm_out_host) && fmt->frames > 0) return fail(FIR_GPU_ERR_INVALID, "null PCM buffer");
	DeviceGuard g(c->device);
	reset_timing(c, true);
	const size_t fb = (size_t) fmt->channels * (fmt->bits / 8);
	const size_t in_bytes = (size_t) (fmt->halo_left + fmt->frames + fmt->halo_right) * fb;
	const size_t out_bytes = (size_t) fmt->frames * fb;
	const bool speculate = !normalize && fmt->frames > 0;
	rc = ensure((void**) &c->d_pcm, &c->pcm_cap, in_bytes + 32);
	if (!rc && speculate) rc = ensure((void**) &c->d_out, &c->out_cap, out_bytes + 32);
	if (rc) return rc;
	rc = pass_begin(c, k, c->d_pcm, fmt, speculate ? 3 : 1);
	if (rc) return rc;
	// Once downloads into pcm_out_host may be in flight, no exit path may return before they
	// have landed (or failed): the caller is free to reuse or free that buffer on return.
	struct DrainOnExit {
		cudaStream_t st;
		~DrainOnExit()
		{
			if (st && cudaStreamSynchronize(st) != cudaSuccess) cudaGetLastError();
		}
	} drain{speculate ? c->d2h_stream : nullptr};
	c->pass.spec_out = speculate ? static_cast<unsigned char*>(pcm_out_host) : nullptr;
	rc = pass_feed_all(c, static_cast<const unsigned char*>(pcm_in_host), fmt, in_bytes);
	if (!rc) rc = pass_end(c);
	if (rc) return rc;
	double pk = 0.0;
	rc = fir_gpu_peak(c, &pk); // waits for the FIR
	if (rc) return rc;
	// ProcessFile.cp:98: normalise when the peak exceeds full scale or -n is given
	const double sc = ((pk > 1.0 || normalize) && pk > 0.0) ? 1.0 / pk : 1.0;
	if (speculate) CU_TRY(cudaStreamSynchronize(c->d2h_stream)); // the scale-1 PCM is in pcm_out_host
	if ((!speculate || sc != 1.0) && fmt->frames > 0) {
		rc = fir_gpu_encode(c, sc, pcm_out_host); // the speculation lost (or was not made): encode for real
		if (rc) return rc;
	}
	if (peak_out) *peak_out = pk;
	if (scale_out) *scale_out = sc;
	return FIR_GPU_OK;
}

// ---- streamed apply: the payload arrives piece by piece (file reads overlap the GPU) ----

int fir_gpu_apply_begin(fir_gpu_ctx* c, const fir_gpu_kernel* k, const fir_gpu_pcm* fmt)
{
	int rc = check_apply_args(c, k, fmt);
	if (rc) return rc;
	DeviceGuard g(c->device);
	reset_timing(c, true);
	const size_t in_bytes =
		(size_t) (fmt->halo_left + fmt->frames + fmt->halo_right) * fmt->channels * (fmt->bits / 8);
	rc = ensure((void**) &c->d_pcm, &c->pcm_cap, in_bytes + 32);
	if (rc) return rc;
	return pass_begin(c, k, c->d_pcm, fmt, 2);
}

int fir_gpu_apply_feed(fir_gpu_ctx* c, const void* pcm_host, size_t bytes)
{
	if (!c || (!pcm_host && bytes)) return fail(FIR_GPU_ERR_INVALID, "null argument");
	if (!c->pass.active || !c->pass.from_host) return fail(FIR_GPU_ERR_STATE, "no streamed apply is open");
	if (c->pass.bytes_fed + bytes > c->pass.bytes_total)
		return fail(FIR_GPU_ERR_INVALID, "more bytes fed than the format announced");
	DeviceGuard g(c->device);
	// the copy of the PREVIOUS feed must be done before the caller reuses that buffer
	// (two alternating host buffers then never wait on the copy in flight)
	CU_TRY(cudaEventSynchronize(c->feed_last));
	int rc = pass_upload(c, static_cast<const unsigned char*>(pcm_host), bytes);
	if (rc) return rc;
	CU_TRY(cudaEventRecord(c->feed_last, c->copy_stream));
	return pass_launch_ready(c);
}

int fir_gpu_apply_end(fir_gpu_ctx* c)
{
	if (!c) return fail(FIR_GPU_ERR_INVALID, "null context");
	if (!c->pass.active) return fail(FIR_GPU_ERR_STATE, "no streamed apply is open");
	DeviceGuard g(c->device);
	// the caller's last buffers are free once the copies are done
	CU_TRY(cudaStreamSynchronize(c->copy_stream));
	return pass_end(c);
}

int fir_gpu_set_progress(fir_gpu_ctx* c, fir_gpu_progress_fn fn, void* user)
{
	if (!c) return fail(FIR_GPU_ERR_INVALID, "null context");
	c->progress = fn;
	c->progress_user = user;
	return FIR_GPU_OK;
}

int fir_gpu_filter_f64(fir_gpu_ctx* c, const fir_gpu_kernel* k, const double* x_host, int64_t frames,
                       int32_t channels, double* y_host)
{
	if (!c || !k || !x_host || !y_host) return fail(FIR_GPU_ERR_INVALID, "null argument");
	if (frames < 1 || channels < 1 || channels > 65535) return fail(FIR_GPU_ERR_INVALID, "bad shape");
	if (k->device != c->device) return fail(FIR_GPU_ERR_STATE, "kernel lives on another device");
	DeviceGuard g(c->device);
	reset_timing(c, true);
	const int64_t H = (k->n_taps - 1) / 2;
	const int64_t x_pitch = x_pitch_for(variant_of(c), frames, k->n_taps);
	c->parked = false;
	c->y_pitch = round_up(frames, 16);
	int rc = ensure((void**) &c->d_x, &c->x_cap, (size_t) x_pitch * channels * sizeof(double));
	if (!rc) rc = ensure((void**) &c->d_y, &c->y_cap, (size_t) c->y_pitch * channels * sizeof(double));
	if (rc) return rc;
	CU_TRY(cudaMemsetAsync(c->d_peak, 0, 8, c->stream));
	CU_TRY(cudaMemsetAsync(c->d_x, 0, (size_t) x_pitch * channels * sizeof(double), c->stream));
	CU_TRY(cudaMemcpy2DAsync(c->d_x + H, (size_t) x_pitch * 8, x_host, (size_t) frames * 8, (size_t) frames * 8,
	                         (size_t) channels, cudaMemcpyHostToDevice, c->stream));
	SPAN_BEGIN(s);
	rc = launch_fir(c, k, c->d_x, x_pitch, channels, c->d_y, c->y_pitch, frames, c->d_peak);
	if (rc) return rc;
	SPAN_END(s);
	c->t_fir.push_back(s);
	CU_TRY(cudaMemcpy2DAsync(y_host, (size_t) frames * 8, c->d_y, (size_t) c->y_pitch * 8, (size_t) frames * 8,
	                         (size_t) channels, cudaMemcpyDeviceToHost, c->stream));
	CU_TRY(cudaStreamSynchronize(c->stream));
	// leave the signal parked so fir_gpu_peak / fir_gpu_parked work on it
	c->fmt = fir_gpu_pcm{frames, channels, 0, 0, 0, 0};
	c->parked = true;
	return FIR_GPU_OK;
}

int fir_gpu_parked(fir_gpu_ctx* c, double* y_host, int64_t frames, int32_t channels)
{
	if (!c || !y_host) return fail(FIR_GPU_ERR_INVALID, "null argument");
	if (!c->parked) return fail(FIR_GPU_ERR_STATE, "no filtered signal is parked on this context");
	if (frames != c->fmt.frames || channels != c->fmt.channels)
		return fail(FIR_GPU_ERR_INVALID, "shape does not match the parked signal");
	DeviceGuard g(c->device);
	if (frames)
		CU_TRY(cudaMemcpy2DAsync(y_host, (size_t) frames * 8, c->d_y, (size_t) c->y_pitch * 8, (size_t) frames * 8,
		                         (size_t) channels, cudaMemcpyDeviceToHost, c->stream));
	CU_TRY(cudaStreamSynchronize(c->stream));
	return FIR_GPU_OK;
}

int fir_gpu_parked_range(fir_gpu_ctx* c, double* y_host, int64_t first_frame, int64_t frames, int32_t channels)
{
	if (!c || !y_host) return fail(FIR_GPU_ERR_INVALID, "null argument");
	if (!c->parked) return fail(FIR_GPU_ERR_STATE, "no filtered signal is parked on this context");
	if (channels != c->fmt.channels || first_frame < 0 || frames < 0 || first_frame + frames > c->fmt.frames)
		return fail(FIR_GPU_ERR_INVALID, "window does not lie inside the parked signal");
	DeviceGuard g(c->device);
	if (frames)
		CU_TRY(cudaMemcpy2DAsync(y_host, (size_t) frames * 8, c->d_y + first_frame, (size_t) c->y_pitch * 8,
		                         (size_t) frames * 8, (size_t) channels, cudaMemcpyDeviceToHost, c->stream));
	CU_TRY(cudaStreamSynchronize(c->stream));
	return FIR_GPU_OK;
}

// -------------------------------------------------------------------- peak

int fir_gpu_peak(fir_gpu_ctx* c, double* peak)
{
	if (!c || !peak) return fail(FIR_GPU_ERR_INVALID, "null argument");
	if (!c->parked) return fail(FIR_GPU_ERR_STATE, "no filtered signal is parked on this context");
	DeviceGuard g(c->device);
	unsigned long long bits = 0;
	CU_TRY(cudaMemcpyAsync(&bits, c->d_peak, 8, cudaMemcpyDeviceToHost, c->stream));
	CU_TRY(cudaStreamSynchronize(c->stream));
	std::memcpy(peak, &bits, 8);
	return FIR_GPU_OK;
}

int fir_gpu_peak_dev(fir_gpu_ctx* c, void** peak_dev)
{
	if (!c || !peak_dev) return fail(FIR_GPU_ERR_INVALID, "null argument");
	if (!c->parked) return fail(FIR_GPU_ERR_STATE, "no filtered signal is parked on this context");
	*peak_dev = c->d_peak;
	return FIR_GPU_OK;
}

int fir_gpu_peak_recompute(fir_gpu_ctx* c, double* peak)
{
	if (!c || !peak) return fail(FIR_GPU_ERR_INVALID, "null argument");
	if (!c->parked) return fail(FIR_GPU_ERR_STATE, "no filtered signal is parked on this context");
	DeviceGuard g(c->device);
	unsigned long long* d_tmp = c->d_peak + 1; // second slot of the 64-byte scratch
	CU_TRY(cudaMemsetAsync(d_tmp, 0, 8, c->stream));
	if (c->fmt.frames > 0) {
		const int64_t pairs = c->fmt.frames / 2;
		unsigned bx = (unsigned) std::min<int64_t>((pairs + 255) / 256, (int64_t) c->sm_count * 8);
		if (bx < 1) bx = 1;
		SPAN_BEGIN(s);
		peak_abs_kernel<<<dim3(bx, (unsigned) c->fmt.channels), 256, 0, c->stream>>>(c->d_y, c->y_pitch,
		                                                                            c->fmt.frames, d_tmp);
		SPAN_END(s);
		c->t_peak.push_back(s);
		c->other_launches++;
		CU_TRY(cudaGetLastError());
	}
	unsigned long long bits = 0;
	CU_TRY(cudaMemcpyAsync(&bits, d_tmp, 8, cudaMemcpyDeviceToHost, c->stream));
	CU_TRY(cudaStreamSynchronize(c->stream));
	std::memcpy(peak, &bits, 8);
	return FIR_GPU_OK;
}

// -------------------------------------------- peak over sample blocks (NCCL)

namespace {

// The few NCCL entry points needed, resolved from libnccl.so.2 at first use: the library is large
// and a single-GPU run (the common case of the CLI) never pays for loading it.
typedef struct ncclComm* nccl_comm_t;
struct NcclApi {
	void* handle = nullptr;
	int (*CommInitAll)(nccl_comm_t*, int, const int*) = nullptr;
	int (*CommDestroy)(nccl_comm_t) = nullptr;
	int (*GroupStart)() = nullptr;
	int (*GroupEnd)() = nullptr;
	int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
	const char* (*GetErrorString)(int) = nullptr;
	std::string error;
};
constexpr int NCCL_UINT64 = 5, NCCL_MAX = 2; // ncclDataType_t / ncclRedOp_t values of nccl.h (stable since 2.0)

NcclApi& nccl_api()
{
	static NcclApi api;
	static std::once_flag once;
	std::call_once(once, [] {
		for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
			api.handle = dlopen(name, RTLD_NOW | RTLD_LOCAL);
			if (api.handle) break;
		}
		if (!api.handle) {
			api.error = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "not found");
			return;
		}
		auto sym = [&](const char* n) {
			void* p = dlsym(api.handle, n);
			if (!p && api.error.empty()) api.error = std::string("libnccl lacks ") + n;
			return p;
		};
		api.CommInitAll = (decltype(api.CommInitAll)) sym("ncclCommInitAll");
		api.CommDestroy = (decltype(api.CommDestroy)) sym("ncclCommDestroy");
		api.GroupStart = (decltype(api.GroupStart)) sym("ncclGroupStart");
		api.GroupEnd = (decltype(api.GroupEnd)) sym("ncclGroupEnd");
		api.AllReduce = (decltype(api.AllReduce)) sym("ncclAllReduce");
		api.GetErrorString = (decltype(api.GetErrorString)) sym("ncclGetErrorString");
	});
	return api;
}

struct CommSet {
	std::vector<int> devices;
	std::vector<nccl_comm_t> comms;
};
std::mutex g_comm_mutex;
std::deque<CommSet> g_comm_sets;  // kept for the life of the process; a deque: references stay valid as it grows

int comm_set_for(fir_gpu_ctx* const* ctxs, int n, CommSet** out)
{
	if (!ctxs || n < 1) return fail(FIR_GPU_ERR_INVALID, "no contexts");
	std::vector<int> devs;
	for (int i = 0; i < n; ++i) {
		if (!ctxs[i]) return fail(FIR_GPU_ERR_INVALID, "null context");
		devs.push_back(ctxs[i]->device);
	}
	for (int i = 0; i < n; ++i)
		for (int j = i + 1; j < n; ++j)
			if (devs[i] == devs[j]) return fail(FIR_GPU_ERR_INVALID, "sample blocks must live on distinct devices");
	NcclApi& api = nccl_api();
	if (!api.error.empty()) return fail(FIR_GPU_ERR_STATE, "NCCL unavailable: " + api.error);
	std::lock_guard<std::mutex> lock(g_comm_mutex);
	for (CommSet& s : g_comm_sets)
		if (s.devices == devs) {
			*out = &s;
			return FIR_GPU_OK;
		}
	CommSet s;
	s.devices = devs;
	s.comms.resize(n);
	const int rc = api.CommInitAll(s.comms.data(), n, devs.data());
	if (rc != 0) return fail(FIR_GPU_ERR_CUDA, std::string("ncclCommInitAll: ") + api.GetErrorString(rc));
	g_comm_sets.push_back(std::move(s));
	*out = &g_comm_sets.back();
	return FIR_GPU_OK;
}

} // namespace

int fir_gpu_comm_prepare(fir_gpu_ctx* const* ctxs, int n)
{
	if (n == 1) return ctxs && ctxs[0] ? FIR_GPU_OK : fail(FIR_GPU_ERR_INVALID, "null context");
	CommSet* s = nullptr;
	return comm_set_for(ctxs, n, &s);
}

int fir_gpu_allreduce_peak(fir_gpu_ctx* const* ctxs, int n, double* peak)
{
	if (!ctxs || n < 1 || !peak) return fail(FIR_GPU_ERR_INVALID, "null argument");
	for (int i = 0; i < n; ++i)
		if (!ctxs[i] || !ctxs[i]->parked) return fail(FIR_GPU_ERR_STATE, "no filtered signal is parked on a context");
	if (n > 1) {
		CommSet* s = nullptr;
		int rc = comm_set_for(ctxs, n, &s);
		if (rc) return rc;
		NcclApi& api = nccl_api();
		std::lock_guard<std::mutex> lock(g_comm_mutex); // one collective of a device set at a time
		// non-negative doubles order like their bit patterns: the scalar the FIR epilogue maintains
		// with atomicMax is reduced as the unsigned integer it is
		rc = api.GroupStart();
		for (int i = 0; i < n && rc == 0; ++i)
			rc = api.AllReduce(ctxs[i]->d_peak, ctxs[i]->d_peak, 1, NCCL_UINT64, NCCL_MAX, s->comms[i], ctxs[i]->stream);
		const int rc_end = api.GroupEnd();
		if (rc == 0) rc = rc_end;
		if (rc != 0) return fail(FIR_GPU_ERR_CUDA, std::string("ncclAllReduce(max): ") + api.GetErrorString(rc));
		for (int i = 1; i < n; ++i) {
			DeviceGuard g(ctxs[i]->device);
			CU_TRY(cudaStreamSynchronize(ctxs[i]->stream));
		}
	}
	return fir_gpu_peak(ctxs[0], peak);
}

// ------------------------------------------------------------------ encode

int fir_gpu_encode_dev(fir_gpu_ctx* c, double scale, void* pcm_dev)
{
	if (!c || !pcm_dev) return fail(FIR_GPU_ERR_INVALID, "null argument");
	if (!c->parked || c->fmt.bits == 0) return fail(FIR_GPU_ERR_STATE, "no PCM-format signal is parked on this context");
	if (!(scale > 0.0) || !std::isfinite(scale)) return fail(FIR_GPU_ERR_INVALID, "scale must be finite and > 0");
	DeviceGuard g(c->device);
	if (c->fmt.frames == 0) return FIR_GPU_OK;
	const double gain = scale * std::ldexp(1.0, c->fmt.bits - 1);
	int rc = FIR_GPU_OK;
	SPAN_BEGIN(s);
	DISPATCH_CODEC(rc, launch_encode, c->fmt.bits, c->fmt.big_endian != 0, c, c->stream, c->d_y, c->y_pitch,
	               c->fmt.frames, c->fmt.channels, gain, (unsigned char*) pcm_dev);
	if (rc) return rc;
	SPAN_END(s);
	c->t_encode.push_back(s);
	return FIR_GPU_OK;
}

// Encode frames [first, first+frames) of the parked signal into the staging buffer and bring
// them to the host, on the high-priority output stream (ordered after the context's stream):
// when another context's FIR fills the GPU, this file's way out does not wait behind it.
static int encode_to_host(fir_gpu_ctx* c, double scale, int64_t first_frame, int64_t frames, void* pcm_host)
{
	const size_t fb = (size_t) c->fmt.channels * (c->fmt.bits / 8);
	int rc = ensure((void**) &c->d_pcm, &c->pcm_cap, (size_t) c->fmt.frames * fb + 32);
	if (rc) return rc;
	if (frames == 0) return FIR_GPU_OK;
	unsigned char* dst = c->d_pcm + (size_t) first_frame * fb;
	const double gain = scale * std::ldexp(1.0, c->fmt.bits - 1);
	CU_TRY(cudaEventRecord(c->enc_ready, c->stream));
	CU_TRY(cudaStreamWaitEvent(c->d2h_stream, c->enc_ready, 0));
	SPAN_BEGIN(se, c->d2h_stream);
	DISPATCH_CODEC(rc, launch_encode, c->fmt.bits, c->fmt.big_endian != 0, c, c->d2h_stream, c->d_y + first_frame,
	               c->y_pitch, frames, c->fmt.channels, gain, dst);
	if (rc) return rc;
	SPAN_END(se, c->d2h_stream);
	c->t_encode.push_back(se);
	SPAN_BEGIN(sc, c->d2h_stream);
	CU_TRY(cudaMemcpyAsync(pcm_host, dst, (size_t) frames * fb, cudaMemcpyDeviceToHost, c->d2h_stream));
	SPAN_END(sc, c->d2h_stream);
	c->t_d2h.push_back(sc);
	CU_TRY(cudaStreamSynchronize(c->d2h_stream));
	return FIR_GPU_OK;
}

int fir_gpu_encode(fir_gpu_ctx* c, double scale, void* pcm_host)
{
	if (!c || !pcm_host) return fail(FIR_GPU_ERR_INVALID, "null argument");
	if (!c->parked || c->fmt.bits == 0) return fail(FIR_GPU_ERR_STATE, "no PCM-format signal is parked on this context");
	if (!(scale > 0.0) || !std::isfinite(scale)) return fail(FIR_GPU_ERR_INVALID, "scale must be finite and > 0");
	DeviceGuard g(c->device);
	return encode_to_host(c, scale, 0, c->fmt.frames, pcm_host);
}

int fir_gpu_encode_range(fir_gpu_ctx* c, double scale, int64_t first_frame, int64_t frames, void* pcm_host)
{
	if (!c || !pcm_host) return fail(FIR_GPU_ERR_INVALID, "null argument");
	if (!c->parked || c->fmt.bits == 0) return fail(FIR_GPU_ERR_STATE, "no PCM-format signal is parked on this context");
	if (!(scale > 0.0) || !std::isfinite(scale)) return fail(FIR_GPU_ERR_INVALID, "scale must be finite and > 0");
	if (first_frame < 0 || frames < 0 || first_frame + frames > c->fmt.frames || (first_frame & 1))
		return fail(FIR_GPU_ERR_INVALID, "range must lie inside the parked signal and start on an even frame");
	DeviceGuard g(c->device);
	return encode_to_host(c, scale, first_frame, frames, pcm_host);
}

// ------------------------------------------------- measurement and synthesis

int fir_gpu_last_timing(fir_gpu_ctx* c, fir_gpu_timing* t)
{
	if (!c || !t) return fail(FIR_GPU_ERR_INVALID, "null argument");
	DeviceGuard g(c->device);
	CU_TRY(cudaStreamSynchronize(c->stream));
	CU_TRY(cudaStreamSynchronize(c->d2h_stream));
	t->h2d_ms = span_ms(c, c->t_h2d);
	t->decode_ms = span_ms(c, c->t_decode);
	t->fir_ms = span_ms(c, c->t_fir);
	t->peak_ms = span_ms(c, c->t_peak);
	t->encode_ms = span_ms(c, c->t_encode);
	t->d2h_ms = span_ms(c, c->t_d2h);
	t->fir_launches = c->fir_launches;
	t->other_launches = c->other_launches;
	return FIR_GPU_OK;
}

int fir_gpu_synth_pcm_dev(fir_gpu_ctx* c, uint64_t seed, int64_t first_frame, int64_t frames, int32_t channels,
                          int32_t bits, int32_t big_endian, int64_t rate, double gain, void* pcm_dev)
{
	if (!c || !pcm_dev) return fail(FIR_GPU_ERR_INVALID, "null argument");
	if (bits != 16 && bits != 24 && bits != 32) return fail(FIR_GPU_ERR_INVALID, "bits must be 16, 24 or 32");
	if (frames < 0 || channels < 1 || rate < 1) return fail(FIR_GPU_ERR_INVALID, "bad shape");
	DeviceGuard g(c->device);
	if (frames == 0) return FIR_GPU_OK;
	const int64_t total = frames * channels;
	const unsigned blocks = (unsigned) std::min<int64_t>((total + 255) / 256, (int64_t) c->sm_count * 16);
	synth_pcm_kernel<<<blocks, 256, 0, c->stream>>>(seed, first_frame, frames, channels, bits, big_endian, rate, gain,
	                                               (unsigned char*) pcm_dev);
	CU_TRY(cudaGetLastError());
	return FIR_GPU_OK;
}

int fir_gpu_copy_probe(fir_gpu_ctx* c, void* host_buf, size_t bytes, int dir, double* ms)
{
	if (!c || !host_buf || !ms) return fail(FIR_GPU_ERR_INVALID, "null argument");
	if (dir != 0 && dir != 1) return fail(FIR_GPU_ERR_INVALID, "dir must be 0 (H2D) or 1 (D2H)");
	DeviceGuard g(c->device);
	int rc = ensure((void**) &c->d_pcm, &c->pcm_cap, bytes + 32);
	if (rc) return rc;
	cudaEvent_t a, b;
	CU_TRY(cudaEventCreate(&a));
	cudaError_t e = cudaEventCreate(&b);
	if (e == cudaSuccess) e = cudaEventRecord(a, c->stream);
	if (e == cudaSuccess)
		e = dir == 0 ? cudaMemcpyAsync(c->d_pcm, host_buf, bytes, cudaMemcpyHostToDevice, c->stream)
		             : cudaMemcpyAsync(host_buf, c->d_pcm, bytes, cudaMemcpyDeviceToHost, c->stream);
	if (e == cudaSuccess) e = cudaEventRecord(b, c->stream);
	if (e == cudaSuccess) e = cudaEventSynchronize(b);
	float f = 0.f;
	if (e == cudaSuccess) e = cudaEventElapsedTime(&f, a, b);
	cudaEventDestroy(a);
	cudaEventDestroy(b);
	CU_TRY(e);
	*ms = (double) f;
	return FIR_GPU_OK;
}

int fir_gpu_reserve(fir_gpu_ctx* c, const fir_gpu_kernel* k, const fir_gpu_pcm* fmt, int host_path)
{
	int rc = check_apply_args(c, k, fmt);
	if (rc) return rc;
	DeviceGuard g(c->device);
	const int ch = fmt->channels;
	const size_t fb = (size_t) ch * (fmt->bits / 8);
	const int64_t y_pitch = round_up(std::max<int64_t>(fmt->frames, 1), 16);
	rc = ensure((void**) &c->d_y, &c->y_cap, (size_t) y_pitch * ch * sizeof(double));
	if (rc) return rc;
	int64_t max_pitch = 0;
	for (int mode = 0; mode < (host_path ? 4 : 1); ++mode)
		for (const auto& [f0, nf] : plan_chunks(c, variant_of(c), fmt->frames, ch, k->n_taps, mode))
			max_pitch = std::max(max_pitch, x_pitch_for(variant_of(c), nf, k->n_taps));
	rc = ensure((void**) &c->d_x, &c->x_cap, (size_t) max_pitch * ch * sizeof(double));
	if (rc || !host_path) return rc;
	rc = ensure((void**) &c->d_pcm, &c->pcm_cap, (size_t) (fmt->halo_left + fmt->frames + fmt->halo_right) * fb + 32);
	if (!rc) rc = ensure((void**) &c->d_out, &c->out_cap, (size_t) fmt->frames * fb + 32);
	return rc;
}

int fir_gpu_set_codec_geometry(fir_gpu_ctx* c, int tile_bytes, int threads, int smem_carveout_pct)
{
	if (!c) return fail(FIR_GPU_ERR_INVALID, "null context");
	if (threads != 128 && threads != 256) return fail(FIR_GPU_ERR_INVALID, "codec threads must be 128 or 256");
	if (tile_bytes != 0 && (tile_bytes < 4096 || tile_bytes > 65536 - 4096))
		return fail(FIR_GPU_ERR_INVALID, "codec tile must be 0 (default: 4096 samples) or 4 KiB .. 60 KiB");
	if (smem_carveout_pct < -1 || smem_carveout_pct > 100)
		return fail(FIR_GPU_ERR_INVALID, "carveout must be -1 (driver's choice) or 0..100 percent");
	c->codec_tile_bytes = tile_bytes;
	c->codec_nt = threads;
	c->codec_carveout = smem_carveout_pct;
	return FIR_GPU_OK;
}

int fir_gpu_test_fail_next_create(int n)
{
	fir_gpu_test_fail_create.store(n < 0 ? 0 : n);
	return FIR_GPU_OK;
}

int fir_gpu_fp64_peak(fir_gpu_ctx* c, int kind, double seconds, double* tflops)
{
	if (!c || !tflops) return fail(FIR_GPU_ERR_INVALID, "null argument");
	if (kind != 0 && kind != 1) return fail(FIR_GPU_ERR_INVALID, "kind must be 0 (DFMA) or 1 (DMMA)");
	DeviceGuard g(c->device);
	const int blocks = c->sm_count * 8, threads = 256;
	cudaEvent_t a, b;
	CU_TRY(cudaEventCreate(&a));
	CU_TRY(cudaEventCreate(&b));
	auto run = [&](int iters) -> double {
		cudaEventRecord(a, c->stream);
		if (kind == 0) dfma_probe_kernel<<<blocks, threads, 0, c->stream>>>(1.0000001, 1e-9, iters, c->d_sink);
		else dmma_probe_kernel<<<blocks, threads, 0, c->stream>>>(1.0000001, 1e-9, iters, c->d_sink);
		cudaEventRecord(b, c->stream);
		cudaEventSynchronize(b);
		float ms = 0.f;
		cudaEventElapsedTime(&ms, a, b);
		return (double) ms * 1e-3;
	};
	// flop per loop trip for the whole grid
	const double per_iter = kind == 0 ? 2.0 * blocks * threads * 16 * 8
	                                  : 512.0 * 8 * 4 * (double) (blocks * (threads / 32));
	run(64); // warm-up
	int iters = 4096;
	double t = run(iters);
	if (seconds > 0 && t > 0) {
		double want = seconds / t * iters;
		if (want > 2.0e9) want = 2.0e9;
		if (want > iters) {
			iters = (int) want;
			t = run(iters);
		}
	}
	cudaEventDestroy(a);
	cudaEventDestroy(b);
	CU_TRY(cudaGetLastError());
	*tflops = per_iter * iters / t * 1e-12;
	return FIR_GPU_OK;
}

} // extern "C"
