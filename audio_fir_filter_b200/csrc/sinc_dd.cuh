// sinc_dd.cuh -- double-double (~106-bit) evaluation of one Blackman windowed-sinc tap.
//
// north_star asks for the kernel "normalised in extended precision"; the reference
// side of that is c_lib's WindowedSinc<float64_t> (ProcessFile.cp:48-50, body absent),
// the oracle restates it in x87 long double (64-bit mantissa).  On the GPU there is no
// long double, so every tap is carried as an unevaluated sum hi+lo of two binary64
// numbers through the whole recipe -- exact angle products, sin(pi x) by argument
// reduction + Taylor series, window, normalising sum, division -- and rounded ONCE at
// the end.  The result is the correctly rounded binary64 tap (the intermediate error
// is ~1e-30 relative): it equals the 50-digit mpmath golden bit for bit
// (tests/golden/taps_*.npz), which the oracle itself only approaches to 1 ulp.
//
// The functions are __host__ __device__ so that tests/ can compile this very file with
// g++ and check it against the golden vectors without a GPU; the product only ever
// calls them from sinc_kernel.cuh's kernels.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define DD_HD __host__ __device__ __forceinline__
#else
#define DD_HD inline
#endif

namespace firgpu {

struct dd {
	double hi, lo;
};

DD_HD dd two_sum(double a, double b)
{
	const double s = a + b;
	const double bb = s - a;
	return {s, (a - (s - bb)) + (b - bb)};
}

DD_HD dd quick_two_sum(double a, double b) // |a| >= |b|
{
	const double s = a + b;
	return {s, b - (s - a)};
}

DD_HD dd two_prod(double a, double b)
{
	const double p = a * b;
	return {p, fma(a, b, -p)};
}

DD_HD dd dd_add(dd a, dd b)
{
	dd s = two_sum(a.hi, b.hi);
	const dd t = two_sum(a.lo, b.lo);
	s.lo += t.hi;
	s = quick_two_sum(s.hi, s.lo);
	s.lo += t.lo;
	return quick_two_sum(s.hi, s.lo);
}

DD_HD dd dd_neg(dd a) { return {-a.hi, -a.lo}; }

DD_HD dd dd_mul(dd a, dd b)
{
	dd p = two_prod(a.hi, b.hi);
	p.lo += a.hi * b.lo + a.lo * b.hi;
	return quick_two_sum(p.hi, p.lo);
}

DD_HD dd dd_mul_d(dd a, double b)
{
	dd p = two_prod(a.hi, b);
	p.lo += a.lo * b;
	return quick_two_sum(p.hi, p.lo);
}

DD_HD dd dd_div(dd a, dd b)
{
	const double q1 = a.hi / b.hi;
	dd r = dd_add(a, dd_neg(dd_mul_d(b, q1)));
	const double q2 = r.hi / b.hi;
	r = dd_add(r, dd_neg(dd_mul_d(b, q2)));
	const double q3 = r.hi / b.hi;
	const dd q = quick_two_sum(q1, q2);
	return dd_add(q, dd{q3, 0.0});
}

DD_HD dd dd_from_ratio(double a, double b) // a / b for exactly representable a, b
{
	return dd_div(dd{a, 0.0}, dd{b, 0.0});
}

// pi, 0.36 and 0.64 to double-double precision (mpmath, 60 digits).
#define DD_PI (firgpu::dd{0x1.921fb54442d18p+1, 0x1.1a62633145c07p-53})
#define DD_036 (firgpu::dd{0x1.70a3d70a3d70ap-2, 0x1.eb851eb851eb8p-57})
#define DD_064 (firgpu::dd{0x1.47ae147ae147bp-1, -0x1.eb851eb851eb8p-57})

// sin(x) / cos(x) for |x| <= pi/4 (+ a hair): Taylor series, Horner in double-double.
// Truncation: x^31/31! and x^30/30! are below 1e-35 at x = pi/4.
DD_HD dd dd_sin_small(dd x)
{
	// (-1)^k / (2k+1)!, k = 1..14
	const double c[14][2] = {
		{-0x1.5555555555555p-3, -0x1.5555555555555p-57},  {0x1.1111111111111p-7, 0x1.1111111111111p-63},
		{-0x1.a01a01a01a01ap-13, -0x1.a01a01a01a01ap-73}, {0x1.71de3a556c734p-19, -0x1.c154f8ddc6c00p-73},
		{-0x1.ae64567f544e4p-26, 0x1.c062e06d1f209p-80},  {0x1.6124613a86d09p-33, 0x1.f28e0cc748ebep-87},
		{-0x1.ae7f3e733b81fp-41, -0x1.1d8656b0ee8cbp-97}, {0x1.952c77030ad4ap-49, 0x1.ac981465ddc6cp-103},
		{-0x1.2f49b46814157p-57, -0x1.2650f61dbdcb4p-112}, {0x1.71b8ef6dcf572p-66, -0x1.d043ae40c4647p-120},
		{-0x1.761b41316381ap-75, 0x1.3423c7d91404fp-130}, {0x1.3f3ccdd165fa9p-84, -0x1.58ddadf344487p-139},
		{-0x1.d1ab1c2dccea3p-94, -0x1.054d0c78aea14p-149}, {0x1.259f98b4358adp-103, 0x1.eaf8c39dd9bc5p-157},
	};
	const dd x2 = dd_mul(x, x);
	dd p = {c[13][0], c[13][1]};
	for (int k = 12; k >= 0; --k) p = dd_add(dd_mul(p, x2), dd{c[k][0], c[k][1]});
	// sin x = x + x * (x^2 * p)
	return dd_add(x, dd_mul(x, dd_mul(x2, p)));
}

DD_HD dd dd_cos_small(dd x)
{
	// (-1)^k / (2k)!, k = 1..14
	const double c[14][2] = {
		{-0x1.0000000000000p-1, 0.0},                     {0x1.5555555555555p-5, 0x1.5555555555555p-59},
		{-0x1.6c16c16c16c17p-10, 0x1.f49f49f49f49fp-65},  {0x1.a01a01a01a01ap-16, 0x1.a01a01a01a01ap-76},
		{-0x1.27e4fb7789f5cp-22, -0x1.cbbc05b4fa99ap-76}, {0x1.1eed8eff8d898p-29, -0x1.2aec959e14c06p-83},
		{-0x1.93974a8c07c9dp-37, -0x1.05d6f8a2efd1fp-92}, {0x1.ae7f3e733b81fp-45, 0x1.1d8656b0ee8cbp-101},
		{-0x1.6827863b97d97p-53, -0x1.eec01221a8b0bp-107}, {0x1.e542ba4020225p-62, 0x1.ea72b4afe3c2fp-120},
		{-0x1.0ce396db7f853p-70, 0x1.aebcdbd20331cp-124}, {0x1.f2cf01972f578p-80, -0x1.9ada5fcc1ab14p-135},
		{-0x1.88e85fc6a4e5ap-89, 0x1.71c37ebd16540p-143}, {0x1.0a18a2635085dp-98, 0x1.b9e2e28e1aa54p-153},
	};
	const dd x2 = dd_mul(x, x);
	dd p = {c[13][0], c[13][1]};
	for (int k = 12; k >= 0; --k) p = dd_add(dd_mul(p, x2), dd{c[k][0], c[k][1]});
	return dd_add(dd{1.0, 0.0}, dd_mul(x2, p));
}

// sin(pi * t) for t = t.hi + t.lo in half-turns, |t| < 2^50.
DD_HD dd dd_sinpi(dd t)
{
	// nearest multiple of 1/2: t = k/2 + r, |r| <= 1/4; the subtraction is exact
	const double k = rint(2.0 * t.hi);
	const dd r = two_sum(t.hi - 0.5 * k, t.lo);
	const dd x = dd_mul(DD_PI, r);
	const long long q = (long long) k & 3; // quadrant (two's complement: right for negative k too)
	switch (q) {
	case 0: return dd_sin_small(x);
	case 1: return dd_cos_small(x);
	case 2: return dd_neg(dd_sin_small(x));
	default: return dd_neg(dd_cos_small(x));
	}
}

// Un-normalised Blackman windowed-sinc low-pass tap i of order M (Smith, DSP Guide
// ch. 16; reference README.md:50,60-62), cutoff fc in cycles/sample:
//   sin(2 pi fc (i-H)) / (i-H)  *  (0.42 - 0.5 cos(2 pi i/M) + 0.08 cos(4 pi i/M)),
// the window in its cancellation-free form u^2 (0.36 + 0.64 u^2), u = sin(pi i/M)
// (0.42 - 0.5 + 0.08 = 0).  Only the left half is evaluated: h[i] == h[M-i] bit for bit.
DD_HD dd dd_lowpass_tap(long long i, long long M, double fc)
{
	const long long H = M / 2;
	if (i > H) i = M - i;
	dd s;
	if (i == H) {
		s = dd_mul_d(DD_PI, 2.0 * fc); // the limit 2 pi fc
	} else {
		const double m = (double) (i - H);
		const dd p = two_prod(2.0 * fc, m); // angle in half-turns, exact
		s = dd_div(dd_sinpi(p), dd{m, 0.0});
	}
	const dd u = dd_sinpi(dd_from_ratio((double) i, (double) M));
	const dd u2 = dd_mul(u, u);
	const dd w = dd_mul(u2, dd_add(DD_036, dd_mul(DD_064, u2)));
	return dd_mul(s, w);
}

// Low-cut tap from the low-pass tap and the double-double sum S of all low-pass
// taps: h = -lp/S, +1 at the centre (spectral inversion), rounded once to binary64.
DD_HD double dd_lowcut_tap(dd lp, dd S, bool centre)
{
	dd q = dd_neg(dd_div(lp, S));
	if (centre) q = dd_add(dd{1.0, 0.0}, q);
	return (q.hi + q.lo) + 0.0; // + 0.0: the exact zero at the window's ends is +0, not -0
}

} // namespace firgpu
