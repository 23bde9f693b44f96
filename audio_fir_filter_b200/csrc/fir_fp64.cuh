// fir_fp64.cuh -- the direct FP64 FIR on the FP64 FMA pipe: the `dfma_*` variants of
// the kernel that replaces apply_filter_range() + WindowedSinc::fms() (reference
// FilterCore.h:20-79).  NOT the default: ncu shows the FMA pipe saturating at ~80 % of its
// nominal rate (29 TFLOP/s), the DMMA formulation in fir_dmma.cuh reaches 99 % (36.9); this
// kernel stays selectable (fir_gpu_set_variant) as the comparison point north_star asks for.
//
//   y[c][n] = sum_{k=0..M} h[k] * xpad[c][n + k],     xpad[c][j + H] = x[c][j]
//
// xpad carries H = M/2 zeros (or real halo samples in sample-block mode) on
// either side, so the prologue / main / epilogue loops of FilterCore.h:57-76
// collapse into one uniform sum -- multiplying by an explicit zero adds an
// exact zero, so the value is that of the clipped sum.
//
// Mapping onto the SM (FP64 FMA pipe bound: 2*taps FLOP per 16 B of output):
//   * a CTA of NT threads owns T_OUT = 16*NT consecutive outputs of one
//     channel; thread t owns the 16 consecutive outputs [16t, 16t+16) and keeps
//     their 16 accumulators in registers for the whole tap loop;
//   * the tap loop runs in tiles of KT taps.  For tile i the CTA needs the
//     KT taps (1-D TMA bulk copy) and the T_OUT+KT samples starting at
//     n0 + i*KT (tiled TMA copy of 128-byte rows = 16 doubles, SWIZZLE_128B),
//     double/triple buffered behind mbarriers so the copies hide under the
//     DFMAs of the previous tile;
//   * inside a tile a thread walks 16 taps per step: it holds a 32-sample
//     window (row t+j, carried in registers from the previous step, and row
//     t+j+1, 8 LDS.128) and issues 256 DFMAs against 16 taps that arrive as 8
//     warp-broadcast LDS.128.  That is 16 shared-memory loads per 256 DFMAs:
//     the FP64 pipe (one warp-DFMA per 2 cycles per SM sub-partition), not the
//     128 B/clk shared-memory port, is the limiter;
//   * thread t reads row t+j+1 while its 7 neighbours in the quarter-warp read
//     the next 7 rows: with the 128-byte TMA swizzle (16-byte chunk c of row r
//     lives at chunk c ^ (r & 7)) those eight 16-byte reads hit eight distinct
//     bank groups -- conflict-free without padding the tile;
//   * each output is ONE sequential FMA chain over k = 0..M in ascending
//     order, so the result does not depend on tiling, grid size, chunking or
//     the number of GPUs (the partition-invariance FilterCore.h gets from
//     per-sample independence);
//   * the epilogue stores the 16 outputs (8 x STG.128 per thread) and folds
//     |y| into the peak: warp shuffle max -> one atomicMax per warp on the bit
//     pattern (non-negative doubles order like unsigned integers).
#pragma once
#include "ptx_sm100.cuh"

namespace firgpu {

constexpr int FIR_R = 16; // outputs per thread == doubles per 128-byte smem row

template <int NT_, int KT_, int STAGES_, int MINB_>
struct FirCfg {
	static constexpr int NT = NT_;         // threads per CTA
	static constexpr int KT = KT_;         // taps per pipeline stage
	static constexpr int STAGES = STAGES_; // smem stages
	static constexpr int MINB = MINB_;     // resident CTAs per SM the register budget targets
	static constexpr int T_OUT = NT * FIR_R;
	static constexpr int ROWS = (T_OUT + KT) / FIR_R; // 128-byte sample rows per stage
	static constexpr int NBOX = (ROWS + 255) / 256;   // TMA box rows <= 256
	static constexpr int BOX_ROWS = ROWS / NBOX;
	static constexpr int SAMPLE_BYTES = ROWS * 128;
	static constexpr int TAP_BYTES = KT * 8;
	static constexpr int STAGE_BYTES = SAMPLE_BYTES + TAP_BYTES;
	static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024; // + manual 1024 B alignment
	static_assert(KT % 32 == 0, "two 16-tap steps per loop trip");
	static_assert(BOX_ROWS * NBOX == ROWS, "sample rows must split into equal TMA boxes");
	static_assert(STAGE_BYTES % 1024 == 0, "every stage must keep the 1024 B swizzle alignment");
	static_assert((BOX_ROWS * 128) % 1024 == 0, "second box must keep the swizzle phase");
};

// One 128-byte row of the swizzled sample tile -> 16 registers.
__device__ __forceinline__ void load_row(double (&dst)[16], const unsigned char* sb, int row)
{
	const unsigned char* rp = sb + row * 128;
	const int sw = (row & 7) << 4;
#pragma unroll
	for (int c = 0; c < 8; ++c) {
		const double2 v = *reinterpret_cast<const double2*>(rp + ((c << 4) ^ sw));
		dst[2 * c] = v.x;
		dst[2 * c + 1] = v.y;
	}
}

// 16 taps x 16 outputs.  lo = samples [0,16) of the window, hi = samples [16,32).
__device__ __forceinline__ void fir_step(double (&acc)[16], const double (&lo)[16], const double (&hi)[16],
                                         const double* __restrict__ tb)
{
#pragma unroll
	for (int kk = 0; kk < 16; kk += 2) {
		const double2 h = *reinterpret_cast<const double2*>(tb + kk); // warp-broadcast LDS.128
#pragma unroll
		for (int r = 0; r < 16; ++r) {
			const int i = r + kk;
			acc[r] = fma(h.x, i < 16 ? lo[i] : hi[i - 16], acc[r]);
		}
#pragma unroll
		for (int r = 0; r < 16; ++r) {
			const int i = r + kk + 1;
			acc[r] = fma(h.y, i < 16 ? lo[i] : hi[i - 16], acc[r]);
		}
	}
}

// grid = (ceil(frames / T_OUT), channels).
//   xmap    : 3-D tensor map over xpad viewed as [channels][rows][16 doubles],
//             box {16, BOX_ROWS, 1}, SWIZZLE_128B, out-of-bounds -> 0
//   taps    : tap array zero-padded to n_ktiles * KT doubles
//   y       : planar output, channel pitch y_pitch doubles, 16-byte aligned
//   peak    : running max of |y| as a uint64 bit pattern (may be null)
template <class Cfg>
__global__ void __launch_bounds__(Cfg::NT, Cfg::MINB)
fir_fp64_kernel(const __grid_constant__ CUtensorMap xmap, const double* __restrict__ taps, int n_ktiles,
                double* __restrict__ y, long long y_pitch, long long frames,
                unsigned long long* __restrict__ peak)
{
	constexpr int KT = Cfg::KT, STAGES = Cfg::STAGES;
	extern __shared__ unsigned char smem_raw[];
	__shared__ __align__(8) unsigned long long full_bar[STAGES];

	const int tid = threadIdx.x;
	const int ch = blockIdx.y;
	const long long n0 = (long long) blockIdx.x * Cfg::T_OUT;
	const int row0 = (int) (n0 / FIR_R); // first sample row of stage 0

	const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
	unsigned char* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

	if (tid == 0) {
		tma_prefetch_desc(&xmap);
#pragma unroll
		for (int s = 0; s < STAGES; ++s) mbar_init(smem_u32(&full_bar[s]), 1);
		fence_barrier_init();
	}
	__syncthreads();

	auto issue = [&](int i) {
		const int s = i % STAGES;
		const uint32_t bar = smem_u32(&full_bar[s]);
		const uint32_t dst = smem_base + s * Cfg::STAGE_BYTES;
		mbar_arrive_expect_tx(bar, Cfg::STAGE_BYTES);
		const int r = row0 + i * (KT / FIR_R);
#pragma unroll
		for (int b = 0; b < Cfg::NBOX; ++b)
			tma_load_3d(dst + b * Cfg::BOX_ROWS * 128, &xmap, 0, r + b * Cfg::BOX_ROWS, ch, bar);
		tma_load_1d(dst + Cfg::SAMPLE_BYTES, taps + (long long) i * KT, Cfg::TAP_BYTES, bar);
	};

	if (tid == 0) {
		for (int i = 0; i < STAGES - 1 && i < n_ktiles; ++i) issue(i);
	}

	double acc[16];
#pragma unroll
	for (int r = 0; r < 16; ++r) acc[r] = 0.0;
	double wa[16], wb[16];

	for (int i = 0; i < n_ktiles; ++i) {
		// Stage (i-1)%STAGES was released by the __syncthreads that ended tile i-1.
		if (tid == 0 && i + STAGES - 1 < n_ktiles) issue(i + STAGES - 1);
		const int s = i % STAGES;
		mbar_wait(smem_u32(&full_bar[s]), (uint32_t) (i / STAGES) & 1u);
		const unsigned char* sb = smem_gen + s * Cfg::STAGE_BYTES;
		const double* tb = reinterpret_cast<const double*>(sb + Cfg::SAMPLE_BYTES);
		// Row tid of tile i holds the samples the thread already carries from the
		// last step of tile i-1; only the very first tile has to fetch it.
		if (i == 0) load_row(wa, sb, tid);
#pragma unroll 1
		for (int j = 0; j < KT / FIR_R; j += 2) {
			load_row(wb, sb, tid + j + 1);
			fir_step(acc, wa, wb, tb + j * FIR_R);
			load_row(wa, sb, tid + j + 2);
			fir_step(acc, wb, wa, tb + (j + 1) * FIR_R);
		}
		__syncthreads();
	}

	// Epilogue: 16 consecutive outputs per thread, guarded at the channel's end.
	const long long n = n0 + (long long) tid * FIR_R;
	double* yp = y + (long long) ch * y_pitch + n;
	double m = 0.0;
	if (n + FIR_R <= frames) {
#pragma unroll
		for (int r = 0; r < 16; r += 2) {
			*reinterpret_cast<double2*>(yp + r) = make_double2(acc[r], acc[r + 1]);
			m = fmax(m, fmax(fabs(acc[r]), fabs(acc[r + 1])));
		}
	} else {
#pragma unroll
		for (int r = 0; r < 16; ++r)
			if (n + r < frames) {
				yp[r] = acc[r];
				m = fmax(m, fabs(acc[r]));
			}
	}
	if (peak != nullptr) {
		m = warp_max(m);
		if ((tid & 31) == 0 && m > 0.0) atomicMax(peak, (unsigned long long) __double_as_longlong(m));
	}
}

} // namespace firgpu
