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
KB_D float kb_sinf(float y) { return kb_sincosf_impl(y, 0); }
KB_D float kb_cosf(float y) { return kb_sincosf_impl(y, 1); }

// tanhf: used only on output paths (TB303 soft clip, SynTHX post-fx), where a last-bit difference
// stays a last-bit difference.  Evaluated in double and rounded once.
KB_D float kb_tanhf(float x) { return (float)tanh((double)x); }
