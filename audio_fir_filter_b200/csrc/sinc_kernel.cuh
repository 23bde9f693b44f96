// sinc_kernel.cuh -- fir_gpu_build_kernel's device side: the Blackman
// windowed-sinc low-cut that the reference gets from c_lib's
// WindowedSinc<float64_t>(fc, bw) + makeLowCut() (ProcessFile.cp:48-50; recipe per
// the reference README.md:50,60-62 -> Smith, DSP Guide ch.16):
//     lp[i] = sin(2 pi fc (i-H)) / (i-H) * (0.42 - 0.5 cos(2 pi i/M) + 0.08 cos(4 pi i/M))
//     h[i]  = -lp[i] / sum(lp),  h[H] += 1
// "Normalised in extended precision": every tap is carried in double-double
// (~106 bit, sinc_dd.cuh) through angle, sine, window, sum and division, and rounded
// to binary64 once.  Three launches: taps + per-block partial sums, the sum, the
// normalised low-cut taps.
#pragma once
#include "ptx_sm100.cuh"
#include "sinc_dd.cuh"

namespace firgpu {

__device__ __forceinline__ dd block_sum_dd(dd v, dd* red)
{
	red[threadIdx.x] = v;
	__syncthreads();
	for (int s = 128; s > 0; s >>= 1) {
		if (threadIdx.x < s) red[threadIdx.x] = dd_add(red[threadIdx.x], red[threadIdx.x + s]);
		__syncthreads();
	}
	return red[0];
}

// Pass 1: lp[i] (double-double) and one double-double partial sum per block.
__global__ void __launch_bounds__(256)
sinc_lowpass_kernel(long long M, double fc, dd* __restrict__ lp, dd* __restrict__ partial)
{
	__shared__ dd red[256];
	const long long i = (long long) blockIdx.x * 256 + threadIdx.x;
	dd v = {0.0, 0.0};
	if (i <= M) {
		v = dd_lowpass_tap(i, M, fc);
		lp[i] = v;
	}
	const dd tot = block_sum_dd(v, red);
	if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}

// Pass 2 (one block): reduce the partials to the double-double sum S.
__global__ void __launch_bounds__(256)
sinc_sum_kernel(const dd* __restrict__ partial, int n_partial, dd* __restrict__ sum_out)
{
	__shared__ dd red[256];
	dd acc = {0.0, 0.0};
	for (int p = threadIdx.x; p < n_partial; p += 256) acc = dd_add(acc, partial[p]);
	const dd tot = block_sum_dd(acc, red);
	if (threadIdx.x == 0) *sum_out = tot;
}

// Pass 3: h[i] = -lp[i]/S (+1 at the centre) for i = 0..M, rounded once.  The caller
// has zero-filled the padded tap array around them.
__global__ void __launch_bounds__(256)
sinc_lowcut_kernel(long long M, const dd* __restrict__ lp, const dd* __restrict__ sum,
                   double* __restrict__ taps, long long n)
{
	const long long i = (long long) blockIdx.x * 256 + threadIdx.x;
	if (i >= n) return;
	taps[i] = dd_lowcut_tap(lp[i], *sum, i == M / 2);
}

// ---- register-resident FP64 throughput probes --------------------------------

// DFMA pipe: 16 independent chains per thread, nothing but DFMA in the loop.
__global__ void __launch_bounds__(256)
dfma_probe_kernel(double a, double b, int iters, double* __restrict__ sink)
{
	double acc[16];
#pragma unroll
	for (int r = 0; r < 16; ++r) acc[r] = (double) (threadIdx.x + r);
#pragma unroll 1
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int u = 0; u < 8; ++u) {
#pragma unroll
			for (int r = 0; r < 16; ++r) acc[r] = fma(acc[r], a, b);
		}
	}
	double s = 0.0;
#pragma unroll
	for (int r = 0; r < 16; ++r) s += acc[r];
	if (s == 123.456) sink[0] = s; // never true: keeps the chains alive
}

// DMMA (legacy tensor path, the only FP64 MMA on sm_100a): mma.sync m8n8k4,
// 8 independent accumulator tiles per warp.
__global__ void __launch_bounds__(256)
dmma_probe_kernel(double a, double b, int iters, double* __restrict__ sink)
{
	double c[8][2];
#pragma unroll
	for (int t = 0; t < 8; ++t) c[t][0] = c[t][1] = (double) (threadIdx.x + t);
#pragma unroll 1
	for (int it = 0; it < iters; ++it) {
#pragma unroll
		for (int u = 0; u < 4; ++u) {
#pragma unroll
			for (int t = 0; t < 8; ++t)
				asm volatile(
					"mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
					: "+d"(c[t][0]), "+d"(c[t][1])
					: "d"(a), "d"(b));
		}
	}
	double s = 0.0;
#pragma unroll
	for (int t = 0; t < 8; ++t) s += c[t][0] + c[t][1];
	if (s == 123.456) sink[0] = s; // never true: keeps the chains alive
}

} // namespace firgpu
