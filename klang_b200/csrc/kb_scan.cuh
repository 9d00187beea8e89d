// klang-b200 — the arithmetic of the tolerance-mode filter scan (KB_FX_TOLERANCE, kb_reverb3.cuh), host + device: the kernel runs it with
// one warp per delay line, tests/host/reverb_scan_check.cpp runs the same functions lane by lane with g++ to measure the error against the
// reference's sequential fp32 filter.
#pragma once
#include "kb_prims.cuh"

#define KB_RV3_SCAN_P 5                       // ticks per lane of the scan (32 x 5 = the 160 ticks of a full chunk)

// ---- tolerance mode: is the re-associated line filter admissible?  Measured on this path (profiles/r02_reverb_tolerance.txt): the fp32
// rounding noise of the TDF-II low-pass (Q = 1/sqrt 2) relative to exact arithmetic, in units of the parity bar, is 0.05 at w = 1.31 rad
// (10 kHz at 48 kHz), 0.13 at 0.65 rad, 2.1 at 0.13 rad: it grows like 1 / w^2.  A scan adds an error of the same size, and the FDN
// recirculates it with a loop gain <= 0.61.  Admitted: poles well inside the unit circle, a2 = r^2 <= 0.5 (w >= ~0.5 rad, 3.8 kHz at 48 kHz).
KB_HD bool kb_rv3_scan_admissible(const KbBiquad& f) { return f.a2 >= 0.f && f.a2 <= 0.5f && f.type == KB_BQ_LPF; }

// ---- tolerance mode: one warp = one line, one chunk.  Lane i owns ticks [5 i, 5 i + 5).
// A tick maps the state s = (z0, z1) to  s' = A s + c x,  A = [[-a1, 1], [-a2, 0]] (from y = b0 x + z0), so a run of 5 ticks from state s ends in
// A^5 s + (the same run from the zero state).  Lane 0 starts from the chunk's true state, the others from zero; an inclusive Kogge-Stone scan
// with the precomputed powers T[j] = (A^5)^(2^j) turns the lane-end states into true ones; every lane then re-runs its ticks from the true
// end state of the lane before it.  FMA is used freely here (this path is not bit-exact by contract).
struct KbRv3ScanCoef { float b0, b1, b2, a1, a2; float T[5][4]; };
KB_HD void kb_rv3_scan_coef(const KbBiquad& f, KbRv3ScanCoef& c) {
	c.b0 = f.b0; c.b1 = f.b1; c.b2 = f.b2; c.a1 = f.a1; c.a2 = f.a2;
	double m[4] = { -(double)f.a1, 1.0, -(double)f.a2, 0.0 }, p[4] = { 1, 0, 0, 1 };
	for (int k = 0; k < KB_RV3_SCAN_P; k++) { const double t[4] = { m[0] * p[0] + m[1] * p[2], m[0] * p[1] + m[1] * p[3], m[2] * p[0] + m[3] * p[2], m[2] * p[1] + m[3] * p[3] }; for (int i = 0; i < 4; i++) p[i] = t[i]; }
	for (int j = 0; j < 5; j++) {
		for (int i = 0; i < 4; i++) c.T[j][i] = (float)p[i];
		const double t[4] = { p[0] * p[0] + p[1] * p[2], p[0] * p[1] + p[1] * p[3], p[2] * p[0] + p[3] * p[2], p[2] * p[1] + p[3] * p[3] };
		for (int i = 0; i < 4; i++) p[i] = t[i];
	}
}
KB_HD float kb_fma(float a, float b, float c) {
#ifdef __CUDA_ARCH__
	return __fmaf_rn(a, b, c);
#else
	return fmaf(a, b, c);
#endif
}
// `cnt` <= KB_RV3_SCAN_P ticks of one lane from state (z0, z1); y (if not null) receives the outputs.  Fully unrolled with a predicate per tick, so x
// and y stay in registers.
KB_HD void kb_rv3_scan_run(const KbRv3ScanCoef& c, const float* x, int cnt, float& z0, float& z1, float* y) {
#ifdef __CUDA_ARCH__
	#pragma unroll
#endif
	for (int j = 0; j < KB_RV3_SCAN_P; j++) {
		if (j < cnt) {
			const float o = kb_fma(c.b0, x[j], z0);
			z0 = kb_fma(c.b1, x[j], kb_fma(-c.a1, o, z1));
			z1 = kb_fma(c.b2, x[j], -c.a2 * o);
			if (y) y[j] = o;
		}
	}
}
