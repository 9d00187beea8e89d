// klang-b200 — device transcendental functions that must agree with the host libm the reference
// is linked against.
//
// Biquad::Filter::set (klang.h:5584-5600) evaluates cosf/sinf per sample when the cutoff moves, and
// b0 = (1-cos0)/2 turns a 1-ulp difference in cosf into ~1e-3 relative at low cutoffs, so "close"
// is not enough: kb_sinf/kb_cosf restate the algorithm of glibc 2.39's sinf/cosf
// (sysdeps/ieee754/flt-32/s_sincosf.h: double-precision minimax polynomials after a fast
// reduction by pi/2) operation by operation.  The restatement is bit-identical to the host libm for
// every float in [0,120) and its negation (exhaustive check on the build host; tests/test_gpu_parity.py
// runs a strided version on the device).  Compiled with -fmad=false: only the KB_MADD sites fuse.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define KB_HD __host__ __device__ __forceinline__
#define KB_D __device__ __forceinline__
#else
#define KB_HD inline
#define KB_D inline
#endif

#ifdef __CUDACC__   // everything below is device code
// glibc selects its FMA build of sinf/cosf (sysdeps/x86_64/fpu/multiarch/s_sinf-fma.c) on every
// x86-64 CPU with FMA+AVX2, i.e. on every B200 host; gcc contracts each `a + b*c` of the generic source
// into one fused operation there.  KB_MADD reproduces exactly that contraction (0 mismatches against the
// host libm over all 1.12e9 floats in [0,120), vs 23 without).  Define KB_LIBM_NO_FMA for a host without FMA.
#ifdef KB_LIBM_NO_FMA
#define KB_MADD(a, b, c) ((a) * (b) + (c))
#else
#define KB_MADD(a, b, c) fma((a), (b), (c))
#endif

KB_D float kb_sincos_poly(double x, double x2, int n, bool neg_cos) {
	// polynomial coefficients of glibc's __sincosf_table (entry 1 negates the cosine polynomial)
	const double c0 = 0x1p0, c1 = -0x1.ffffffd0c621cp-2, c2 = 0x1.55553e1068f19p-5, c3 = -0x1.6c087e89a359dp-10, c4 = 0x1.99343027bf8c3p-16;
	const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
	if ((n & 1) == 0) {
		const double x3 = x * x2;
		const double t1 = KB_MADD(x2, s3, s2);
		const double x7 = x3 * x2;
		const double s = KB_MADD(x3, s1, x);
		return (float)KB_MADD(x7, t1, s);
	} else {
		const double sg = neg_cos ? -1.0 : 1.0;
		const double x4 = x2 * x2;
		const double t2 = KB_MADD(x2, sg * c4, sg * c3);
		const double t1 = KB_MADD(x2, sg * c1, sg * c0);
		const double x6 = x4 * x2;
		const double c = KB_MADD(x4, sg * c2, t1);
		return (float)KB_MADD(x6, t2, c);
	}
}

KB_D uint32_t kb_abstop12(float x) { return (__float_as_uint(x) >> 20) & 0x7ff; }

// glibc sinf / cosf for |y| < 120 (the reduce_fast branch); larger arguments never occur on the hot
// path (w = f * 2pi/fs <= pi) and fall back to the double-precision routine.
KB_D float kb_sincosf_impl(float y, int cosine) {
	double x = (double)y;
	if (kb_abstop12(y) < 0x3f4 /* abstop12(pi/4) */) {
		const double x2 = x * x;
		if (kb_abstop12(y) < 0x398 /* abstop12(2^-12) */) return cosine ? 1.0f : y;
		return kb_sincos_poly(x, x2, cosine, false);
	} else if (kb_abstop12(y) < 0x42f /* abstop12(120) */) {
		const double hpi_inv = 0x1.45F306DC9C883p+23, hpi = 0x1.921FB54442D18p0;
		const double r = x * hpi_inv;
		const int n = ((int32_t)r + 0x800000) >> 24;
		x = KB_MADD(-(double)n, hpi, x);
		const double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
		return kb_sincos_poly(x * s, x * x, n ^ cosine, (n & 2) != 0);
	}
	return cosine ? (float)cos(x) : (float)sin(x);
}
// sinf and cosf of the same argument with one shared reduction (each result identical to the separate calls)
KB_D void kb_sincosf(float y, float& sn, float& cs) {
	double x = (double)y;
	if (kb_abstop12(y) < 0x3f4) {
		const double x2 = x * x;
		if (kb_abstop12(y) < 0x398) { sn = y; cs = 1.0f; return; }
		sn = kb_sincos_poly(x, x2, 0, false);
		cs = kb_sincos_poly(x, x2, 1, false);
	} else if (kb_abstop12(y) < 0x42f) {
		const double hpi_inv = 0x1.45F306DC9C883p+23, hpi = 0x1.921FB54442D18p0;
		const double r = x * hpi_inv;
		const int n = ((int32_t)r + 0x800000) >> 24;
		x = KB_MADD(-(double)n, hpi, x);
		const double s = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
		const double xs = x * s, x2 = x * x;
		sn = kb_sincos_poly(xs, x2, n, (n & 2) != 0);
		cs = kb_sincos_poly(xs, x2, n ^ 1, (n & 2) != 0);
	} else { sn = (float)sin(x); cs = (float)cos(x); }
}
KB_D float kb_sinf(float y) { return kb_sincosf_impl(y, 0); }
KB_D float kb_cosf(float y) { return kb_sincosf_impl(y, 1); }

// expf (TB303.k's Filter::set evaluates `exp()` per sample; OnePole::set does when a program sets it per sample): glibc 2.39's expf
// (sysdeps/ieee754/flt-32/e_expf.c — the ARM optimized-routines algorithm: x*N/ln2 = k + r, 2^(k/N) from a 32-entry table, a cubic in r, all in
// double) restated operation by operation.  The table is 2^(i/32) with the exponent pre-subtracted (bits(2^(i/32)) - (i << 47)), regenerated from
// exact integer arithmetic.  glibc's FMA build (e_expf-fma.c, selected on every x86-64 host with FMA + AVX2) contracts four sites — `r = z - kd`
// with z = InvLn2N * x, and the three polynomial steps; with exactly these the restatement is bit-identical to the build host's libm for ALL
// 2^32 floats (without the first one, two inputs differ by an ulp).  KB_LIBM_NO_FMA for a host without FMA.
__device__ __constant__ unsigned long long kb_exp2f_tab[32] = {
0x3ff0000000000000ull,
0x3fefd9b0d3158574ull,
0x3fefb5586cf9890full,
0x3fef9301d0125b51ull,
0x3fef72b83c7d517bull,
0x3fef54873168b9aaull,
0x3fef387a6e756238ull,
0x3fef1e9df51fdee1ull,
0x3fef06fe0a31b715ull,
0x3feef1a7373aa9cbull,
0x3feedea64c123422ull,
0x3feece086061892dull,
0x3feebfdad5362a27ull,
0x3feeb42b569d4f82ull,
0x3feeab07dd485429ull,
0x3feea47eb03a5585ull,
0x3feea09e667f3bcdull,
0x3fee9f75e8ec5f74ull,
0x3feea11473eb0187ull,
0x3feea589994cce13ull,
0x3feeace5422aa0dbull,
0x3feeb737b0cdc5e5ull,
0x3feec49182a3f090ull,
0x3feed503b23e255dull,
0x3feee89f995ad3adull,
0x3feeff76f2fb5e47ull,
0x3fef199bdd85529cull,
0x3fef3720dcef9069ull,
0x3fef5818dcfba487ull,
0x3fef7c97337b9b5full,
0x3fefa4afa2a490daull,
0x3fefd0765b6e4540ull
};
KB_D float kb_expf(float x) {
	const double InvLn2N = 0x1.71547652b82fep+0 * 32, SHIFT = 0x1.8p+52;
	const double C0 = 0x1.c6af84b912394p-5 / 32 / 32 / 32, C1 = 0x1.ebfce50fac4f3p-3 / 32 / 32, C2 = 0x1.62e42ff0c52d6p-1 / 32;
	const double xd = (double)x;
	const uint32_t abstop = kb_abstop12(x);
	if (abstop >= 0x42b /* abstop12(88.0f) */) {
		if (__float_as_uint(x) == 0xff800000u) return 0.0f;               // -inf
		if (abstop >= 0x7f8) return x + x;                                // +inf, nan
		if (x > 0x1.62e42ep6f) return __uint_as_float(0x7f800000u);       // overflow
		if (x < -0x1.9fe368p6f) return 0.0f;                              // underflow
	}
	const double z = InvLn2N * xd;
	double kd = z + SHIFT;
	const unsigned long long ki = (unsigned long long)__double_as_longlong(kd);
	kd -= SHIFT;
	const double r = KB_MADD(InvLn2N, xd, -kd);
	unsigned long long t = kb_exp2f_tab[ki % 32];
	t += ki << (52 - 5);
	const double s = __longlong_as_double((long long)t);
	const double p = KB_MADD(C0, r, C1);
	const double r2 = r * r;
	double y = KB_MADD(C2, r, 1.0);
	y = KB_MADD(p, r2, y);
	y = y * s;
	return (float)y;
}

// tanhf (TB303 soft clip, SynTHX post-fx): glibc 2.39 still ships the FDLIBM float routines
// (sysdeps/ieee754/flt-32/s_tanhf.c on top of s_expm1f.c, no FMA variant), restated here operation by operation in
// fp32; bit-identical to the host libm for every finite float (exhaustive check on the build host).
KB_D float kb_expm1f(float x) {
	const float one = 1.0f, huge = 1.0e+30f, tiny = 1.0e-30f, o_threshold = 8.8721679688e+01f,
	            ln2_hi = 6.9313812256e-01f, ln2_lo = 9.0580006145e-06f, invln2 = 1.4426950216e+00f,
	            Q1 = -3.3333335072e-02f, Q2 = 1.5873016091e-03f, Q3 = -7.9365076090e-05f, Q4 = 4.0082177293e-06f, Q5 = -2.0109921195e-07f;
	float y, hi, lo, c = 0.f, t, e, hxs, hfx, r1;
	int k;
	uint32_t hx = __float_as_uint(x);
	const uint32_t xsb = hx & 0x80000000u;
	hx &= 0x7fffffffu;
	if (hx >= 0x4195b844u) {                  /* |x| >= 27 ln2 */
		if (hx >= 0x42b17218u) {
			if (hx > 0x7f800000u) return x + x;
			if (hx == 0x7f800000u) return (xsb == 0) ? x : -1.0f;
			if (x > o_threshold) return huge * huge;
		}
		if (xsb != 0) return tiny - one;
	}
	if (hx > 0x3eb17218u) {                   /* |x| > 0.5 ln2 */
		if (hx < 0x3F851592u) {               /* |x| < 1.5 ln2 */
			if (xsb == 0) { hi = x - ln2_hi; lo = ln2_lo; k = 1; }
			else { hi = x + ln2_hi; lo = -ln2_lo; k = -1; }
		} else {
			k = (int)(invln2 * x + ((xsb == 0) ? 0.5f : -0.5f));
			t = (float)k;
			hi = x - t * ln2_hi;
			lo = t * ln2_lo;
		}
		x = hi - lo;
		c = (hi - x) - lo;
	} else if (hx < 0x33000000u) {            /* |x| < 2^-25 */
		t = huge + x;
		return x - (t - (huge + x));
	} else k = 0;
	hfx = 0.5f * x;
	hxs = x * hfx;
	r1 = one + hxs * (Q1 + hxs * (Q2 + hxs * (Q3 + hxs * (Q4 + hxs * Q5))));
	t = 3.0f - r1 * hfx;
	e = hxs * ((r1 - t) / (6.0f - x * t));
	if (k == 0) return x - (x * e - hxs);
	e = (x * (e - c) - c);
	e -= hxs;
	if (k == -1) return 0.5f * (x - e) - 0.5f;
	if (k == 1) { if (x < -0.25f) return -2.0f * (e - (x + 0.5f)); else return one + 2.0f * (x - e); }
	if (k <= -2 || k > 56) {
		y = one - (e - x);
		y = __uint_as_float(__float_as_uint(y) + ((uint32_t)k << 23));
		return y - one;
	}
	if (k < 23) {
		t = __uint_as_float(0x3f800000u - (0x1000000u >> k));
		y = t - (e - x);
		y = __uint_as_float(__float_as_uint(y) + ((uint32_t)k << 23));
	} else {
		t = __uint_as_float((uint32_t)(0x7f - k) << 23);
		y = x - (e + t);
		y += one;
		y = __uint_as_float(__float_as_uint(y) + ((uint32_t)k << 23));
	}
	return y;
}
KB_D float kb_tanhf(float x) {
	const float one = 1.0f, tiny = 1.0e-30f;
	float t, z;
	const int jx = (int)__float_as_uint(x);
	const int ix = jx & 0x7fffffff;
	if (ix >= 0x7f800000) return (jx >= 0) ? one / x + one : one / x - one;
	if (ix < 0x41b00000) {                    /* |x| < 22 */
		if (ix == 0) return x;
		if (ix < 0x24000000) return x * (one + x);
		if (ix >= 0x3f800000) { t = kb_expm1f(2.0f * fabsf(x)); z = one - 2.0f / (t + 2.0f); }
		else { t = kb_expm1f(-2.0f * fabsf(x)); z = -t / (t + 2.0f); }
	} else z = one - tiny;
	return (jx >= 0) ? z : -z;
}
#endif  // __CUDACC__
