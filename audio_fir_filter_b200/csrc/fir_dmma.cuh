// fir_dmma.cuh -- the direct FP64 FIR as a blocked-Toeplitz contraction on the
// FP64 tensor path (DMMA.8x8x4, the only FP64 MMA sm_100a has; tcgen05 has no
// FP64 kind).  Same arithmetic contract as fir_fp64.cuh (reference
// FilterCore.h:20-79):
//
//   y[c][n] = sum_{k=0..M} h[k] * xpad[c][n + k]
//
// Why a second kernel: ncu on the DFMA kernel (profiles/r1_fir_dfma_*.txt) shows
// the FP64 FMA pipe saturated (stall = math_pipe_throttle) at ~80 % of its
// nominal rate, and a register-resident DFMA probe tops out at 91 %, while the
// DMMA probe reaches 99.7 % of 148 SM x 64 FMA/clk.  One DMMA does 256 MACs for
// one issue slot and 4 operand registers, so the issue port and the register
// file stop being the limiter.  Measured with this kernel: DMMA pipe 99.4 % active,
// 36.9 TFLOP/s on config 2 (profiles/r1_fir_dmma_cfg2_ncu.txt).
//
// Blocking.  A warp owns T consecutive tiles of 64 outputs.  For tile base b
// and tap step s (8 taps per step, two MMAs "even"/"odd"):
//
//   D[r][i] = y[b + 8r + i]                           r, i = 0..7   (accumulator)
//   A[r][q] = xpad[b + 8s + 8r + 2q + p]              q = 0..3      (data,  p = 0 even / 1 odd)
//   B[q][i] = h[8s + 2q + p - i]                                    (taps, Toeplitz block)
//
// so that A[r][q] * B[q][i] = h[k] * xpad[(b + 8r + i) + k] with k = 8s+2q+p-i:
// every MAC is a wanted one except k < 0 or k > M, where the zero-padded tap
// array contributes an exact zero ((M+1)/(M+8) of the MACs are useful).
//
// Fragments.  mma.m8n8k4 wants A[r][q] in lane 4r+q: that is
// xpad[b + 8s + 2*lane + p], i.e. ONE conflict-free LDS.128 per lane fetches the
// A operand of both MMAs of a step from a plainly laid out sample tile.  B[q][i]
// sits in lane 4i+q: two LDS.64 of h[8s + 2q - i (+1)], ten distinct
// consecutive words per half-warp, conflict-free.  The accumulator lane owns
// y[b + 2*lane], y[b + 2*lane + 1]: the epilogue is one coalesced STG.128.
//
// Register ring.  The A fragment of tile t at step s is G(8t + s), G(v) =
// xpad[b0 + 8v + 2*lane ..+1]: tile t+1 needs now what tile t needs 8 steps
// later.  So each step LOADS only G(8(T-1) + s), for the last tile, and keeps it
// in a ring of 8(T-1) fragments for the other tiles: 1 LDS.128 + 2 LDS.64 per 2T
// MMAs.  With T = 3 that is 8 shared-memory wavefronts per 96 DMMA-pipe cycles
// per sub-partition, ~1/3 of the 128 B/clk port.
//
// Summation order.  Output n accumulates its taps in ascending groups
// {8s + 2q + p - (n mod 8)}; the grouping depends only on n mod 8, and tile
// bases, chunk starts and sample-block starts are all multiples of 16, so the
// bits do not depend on tiling, chunking or the number of GPUs.
//
// Staging: per tap tile of KT taps the CTA needs T_OUT + KT samples and KT + 16
// taps, each ONE 1-D TMA bulk copy (cp.async.bulk -> UBLKCP) behind an mbarrier,
// STAGES deep.
#pragma once
#include "ptx_sm100.cuh"

namespace firgpu {

template <int NT_, int T_, int KT_, int STAGES_, int MINB_>
struct DmmaCfg {
	static constexpr int NT = NT_;         // threads per CTA
	static constexpr int T = T_;           // 64-output tiles per warp
	static constexpr int KT = KT_;         // taps per pipeline stage
	static constexpr int STAGES = STAGES_;
	static constexpr int MINB = MINB_;
	static constexpr int WARPS = NT / 32;
	static constexpr int T_OUT = WARPS * T * 64;
	static constexpr int TAP_PAD = 8;                       // zeros in front of h[0]
	static constexpr int SAMPLE_BYTES = (T_OUT + KT) * 8;
	static constexpr int TAP_BYTES = (KT + 16) * 8;
	static constexpr int STAGE_BYTES = SAMPLE_BYTES + TAP_BYTES;
	static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 128;
	static_assert(KT % 64 == 0, "8 ring positions x 8 taps per unrolled trip");
	static_assert(STAGE_BYTES % 16 == 0, "bulk copies need 16-byte granularity");
};

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b)
{
	asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
	             : "+d"(c[0]), "+d"(c[1])
	             : "d"(a), "d"(b));
}

// grid = (ceil(frames / T_OUT), channels).
//   xpad    : planar zero-padded input, channel pitch x_pitch doubles; readable up to
//             gridDim.x*T_OUT + n_ktiles*KT per channel
//   tpad    : TAP_PAD zeros, the M+1 taps, zeros up to n_ktiles*KT + 16 doubles
//   n_steps : 8-tap steps that touch a real tap, floor((M+7)/8) + 1 <= n_ktiles*KT/8; the
//             last tap tile stops there instead of multiplying its zero padding
//   y, peak : as in fir_fp64_kernel
template <class Cfg>
__global__ void __launch_bounds__(Cfg::NT, Cfg::MINB)
fir_dmma_kernel(const double* __restrict__ xpad, long long x_pitch, const double* __restrict__ tpad, int n_ktiles,
                int n_steps, double* __restrict__ y, long long y_pitch, long long frames,
                unsigned long long* __restrict__ peak)
{
	constexpr int T = Cfg::T, KT = Cfg::KT, STAGES = Cfg::STAGES;
	constexpr int RING = 8 * (T - 1);
	extern __shared__ __align__(128) unsigned char smem_raw[];
	__shared__ __align__(8) unsigned long long full_bar[STAGES];

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int ch = blockIdx.y;
	const long long n0 = (long long) blockIdx.x * Cfg::T_OUT;
	const double* xsrc = xpad + (long long) ch * x_pitch + n0;

	const uint32_t smem_base = (smem_u32(smem_raw) + 127u) & ~127u;
	unsigned char* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

	if (tid == 0) {
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
		tma_load_1d(dst, xsrc + (long long) i * KT, Cfg::SAMPLE_BYTES, bar);
		tma_load_1d(dst + Cfg::SAMPLE_BYTES, tpad + (long long) i * KT, Cfg::TAP_BYTES, bar);
	};
	if (tid == 0) {
		for (int i = 0; i < STAGES - 1 && i < n_ktiles; ++i) issue(i);
	}

	double acc[T][2];
#pragma unroll
	for (int t = 0; t < T; ++t) acc[t][0] = acc[t][1] = 0.0;
	double2 ring[RING > 0 ? RING : 1];

	// per-lane offsets (in doubles) inside a stage
	const int a_off = warp * (T * 64) + 2 * lane;                    // tile 0 of this warp, step 0
	const int b_off = Cfg::TAP_PAD + 2 * (lane & 3) - (lane >> 2);   // h[2q - i] of step 0

	for (int i = 0; i < n_ktiles; ++i) {
		if (tid == 0 && i + STAGES - 1 < n_ktiles) issue(i + STAGES - 1);
		const int s = i % STAGES;
		mbar_wait(smem_u32(&full_bar[s]), (uint32_t) (i / STAGES) & 1u);
		const double* xs = reinterpret_cast<const double*>(smem_gen + s * Cfg::STAGE_BYTES);
		const double* ts = xs + (Cfg::T_OUT + KT);
		if (i == 0) {
			// preload: delay group d, position j must hold L(j - 8(d+1)) =
			// G(8(T-2-d) + j), the fragments the first T-1 tiles need before the
			// last tile's loads reach them
#pragma unroll
			for (int d = 0; d < T - 1; ++d)
#pragma unroll
				for (int j = 0; j < 8; ++j)
					ring[8 * d + j] = *reinterpret_cast<const double2*>(xs + a_off + 8 * (8 * (T - 2 - d) + j));
		}
		const double* ap = xs + a_off + 64 * (T - 1);
		const double* bp = ts + b_off;
		const int steps = min(KT / 8, n_steps - i * (KT / 8)); // < KT/8 only in the last tile
		// One trip = 8 steps (one full turn of the ring), unguarded so that the loads of all 8
		// steps can be hoisted.  Every tile but the last runs a compile-time number of trips.
		auto trip = [&](int j0) {
#pragma unroll
			for (int j = 0; j < 8; ++j) {
				const double2 g = *reinterpret_cast<const double2*>(ap + 8 * (j0 + j));
				const double be = bp[8 * (j0 + j)];
				const double bo = bp[8 * (j0 + j) + 1];
				// tile t uses the fragment loaded 8(T-1-t) steps ago: ring slot
				// 8(T-2-t) + j holds it (slot j of delay group T-2-t)
#pragma unroll
				for (int t = 0; t < T - 1; ++t) dmma(acc[t], ring[8 * (T - 2 - t) + j].x, be);
				dmma(acc[T - 1], g.x, be);
#pragma unroll
				for (int t = 0; t < T - 1; ++t) dmma(acc[t], ring[8 * (T - 2 - t) + j].y, bo);
				dmma(acc[T - 1], g.y, bo);
				// age the ring position j by one delay group
#pragma unroll
				for (int d = T - 2; d > 0; --d) ring[8 * d + j] = ring[8 * (d - 1) + j];
				if (T > 1) ring[j] = g;
			}
		};
		const int full = steps & ~7;
		if (i + 1 < n_ktiles) {
#pragma unroll
			for (int j0 = 0; j0 < KT / 8; j0 += 8) trip(j0);
		} else {
#pragma unroll 1
			for (int j0 = 0; j0 < full; j0 += 8) trip(j0);
		}
		// The up to 7 steps left in the LAST tile (the taps end there): no ring any more, every
		// tile's fragment comes straight from the sample tile.
#pragma unroll 1
		for (int j = full; j < steps; ++j) {
			const double be = bp[8 * j];
			const double bo = bp[8 * j + 1];
#pragma unroll
			for (int t = 0; t < T; ++t) {
				const double2 g = *reinterpret_cast<const double2*>(xs + a_off + 64 * t + 8 * j);
				dmma(acc[t], g.x, be);
				dmma(acc[t], g.y, bo);
			}
		}
		__syncthreads();
	}

	// Epilogue: lane owns y[b + 2*lane], y[b + 2*lane + 1] of each of its tiles.
	double m = 0.0;
#pragma unroll
	for (int t = 0; t < T; ++t) {
		const long long n = n0 + warp * (T * 64) + t * 64 + 2 * lane;
		double* yp = y + (long long) ch * y_pitch + n;
		if (n + 1 < frames) {
			*reinterpret_cast<double2*>(yp) = make_double2(acc[t][0], acc[t][1]);
			m = fmax(m, fmax(fabs(acc[t][0]), fabs(acc[t][1])));
		} else if (n < frames) {
			yp[0] = acc[t][0];
			m = fmax(m, fabs(acc[t][0]));
		}
	}
	if (peak != nullptr) {
		m = warp_max(m);
		if (lane == 0 && m > 0.0) atomicMax(peak, (unsigned long long) __double_as_longlong(m));
	}
}

} // namespace firgpu
