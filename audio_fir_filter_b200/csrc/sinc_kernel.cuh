// sinc_kernel.cuh -- fir_gpu_build_kernel's device side: the Blackman
// windowed-sinc low-cut that the reference gets from c_lib's
// WindowedSinc<float64_t>(fc, bw) + makeLowCut() (ProcessFile.cp:48-50; recipe per
// the reference README.md:50,60-62 -> Smith, DSP Guide ch.16):
//     lp[i] = sin(2 pi fc (i-H)) / (i-H) * (0.42 - 0.5 cos(2 pi i/M) + 0.08 cos(4 pi i/M))
//             (the window is evaluated in its cancellation-free form, see lowpass_tap)
//     h[i]  = -lp[i] / sum(lp),  h[H] += 1
// "Normalised in extended precision": the sum runs in double-double (~106 bit),
// the angles are reduced exactly (products and quotients keep their FMA
// remainders and feed sinpi/cospi, so there is no 2*pi*x range-reduction error),
// and the division by the double-double sum is corrected with one Newton step.
#pragma once
#include "ptx_sm100.cuh"

namespace firgpu {

struct dd {
	double hi, lo;
};

__device__ __forceinline__ dd two_sum(double a, double b)
{
	const double s = __dadd_rn(a, b);
	const double bb = __dsub_rn(s, a);
	const double e = __dadd_rn(__dsub_rn(a, __dsub_rn(s, bb)), __dsub_rn(b, bb));
	return {s, e};
}

__device__ __forceinline__ dd dd_add(dd a, dd b)
{
	dd s = two_sum(a.hi, b.hi);
	const dd t = two_sum(a.lo, b.lo);
	s.lo = __dadd_rn(s.lo, t.hi);
	s = two_sum(s.hi, s.lo); // renormalise
	s.lo = __dadd_rn(s.lo, t.lo);
	return two_sum(s.hi, s.lo);
}

constexpr double PI_HI = 3.141592653589793116e+00; // pi rounded to binary64
constexpr double PI_LO = 1.224646799147353207e-16; // pi - PI_HI

// sin(pi * (p_hi + p_lo)) and cos(pi * (p_hi + p_lo)) for |p_lo| << 1:
// first-order correction around the binary64 angle.
__device__ __forceinline__ double sinpi_dd(double p_hi, double p_lo)
{
	double s, c;
	sincospi(p_hi, &s, &c);
	return fma(__dmul_rn(PI_HI, p_lo), c, s);
}

__device__ __forceinline__ double cospi_dd(double p_hi, double p_lo)
{
	double s, c;
	sincospi(p_hi, &s, &c);
	return fma(-__dmul_rn(PI_HI, p_lo), s, c);
}

// Un-normalised low-pass tap i (binary64).
__device__ __forceinline__ double lowpass_tap(long long i, long long M, double fc)
{
	const long long H = M / 2;
	if (i > H) i = M - i; // evaluate the left half only: h[i] == h[M-i] bit for bit
	const double m = (double) (i - H);
	double s;
	if (i == H) {
		// 2*pi*fc, rounded once
		s = fma(2.0 * fc, PI_HI, 2.0 * fc * PI_LO);
	} else {
		const double two_fc = 2.0 * fc;                 // exact
		const double p_hi = __dmul_rn(two_fc, m);       // angle in half-turns
		const double p_lo = fma(two_fc, m, -p_hi);      // exact remainder
		s = __ddiv_rn(sinpi_dd(p_hi, p_lo), m);
	}
	// Blackman window without cancellation: with u = sin(pi i / M),
	//   0.42 - 0.5 cos(2 pi i/M) + 0.08 cos(4 pi i/M) = u^2 (0.36 + 0.64 u^2)
	// (0.42 - 0.5 + 0.08 = 0), so the small taps at the ends keep full relative
	// accuracy instead of the ~1e-17 absolute error of the three-term form.
	const double di = (double) i, dM = (double) M;
	const double t_hi = __ddiv_rn(di, dM);              // i / M half-turns
	const double t_lo = __ddiv_rn(fma(-t_hi, dM, di), dM);
	const double u = sinpi_dd(t_hi, t_lo);
	const double u2 = __dmul_rn(u, u);
	const double win = __dmul_rn(u2, fma(0.64, u2, 0.36));
	return __dmul_rn(s, win);
}

// Pass 1: lp[i] and one double-double partial sum per block.
__global__ void __launch_bounds__(256)
sinc_lowpass_kernel(long long M, double fc, double* __restrict__ lp, dd* __restrict__ partial)
{
	__shared__ dd red[256];
	const long long i = (long long) blockIdx.x * 256 + threadIdx.x;
	double v = 0.0;
	if (i <= M) {
		v = lowpass_tap(i, M, fc);
		lp[i] = v;
	}
	red[threadIdx.x] = {v, 0.0};
	__syncthreads();
	for (int s = 128; s > 0; s >>= 1) {
		if (threadIdx.x < s) red[threadIdx.x] = dd_add(red[threadIdx.x], red[threadIdx.x + s]);
		__syncthreads();
	}
	if (threadIdx.x == 0) partial[blockIdx.x] = red[0];
}

// Pass 2 (one block): reduce the partials to the double-double sum S.
__global__ void __launch_bounds__(256)
sinc_sum_kernel(const dd* __restrict__ partial, int n_partial, dd* __restrict__ sum_out)
{
	__shared__ dd red[256];
	dd acc = {0.0, 0.0};
	for (int p = threadIdx.x; p < n_partial; p += 256) acc = dd_add(acc, partial[p]);
	red[threadIdx.x] = acc;
	__syncthreads();
	for (int s = 128; s > 0; s >>= 1) {
		if (threadIdx.x < s) red[threadIdx.x] = dd_add(red[threadIdx.x], red[threadIdx.x + s]);
		__syncthreads();
	}
	if (threadIdx.x == 0) *sum_out = red[0];
}

// Pass 3: h[i] = -lp[i]/S (+1 at the centre) for i = 0..M.  The caller has
// zero-filled the padded tap array around them.
__global__ void __launch_bounds__(256)
sinc_lowcut_kernel(long long M, const double* __restrict__ lp, const dd* __restrict__ sum,
                   double* __restrict__ taps, long long n)
{
	const long long i = (long long) blockIdx.x * 256 + threadIdx.x;
	if (i >= n) return;
	const dd S = *sum;
	const double h = lp[i];
	// q = h / (S.hi + S.lo) as a double-double quotient
	const double q0 = __ddiv_rn(h, S.hi);
	const double r = __dsub_rn(fma(-q0, S.hi, h), __dmul_rn(q0, S.lo));
	const double q1 = __ddiv_rn(r, S.hi);
	if (i == M / 2) {
		dd one = two_sum(1.0, -q0);
		taps[i] = __dadd_rn(one.hi, __dsub_rn(one.lo, q1));
	} else {
		taps[i] = -__dadd_rn(q0, q1);
	}
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
	if (s == 123.456) sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
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
	if (s == 123.456) sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

} // namespace firgpu
