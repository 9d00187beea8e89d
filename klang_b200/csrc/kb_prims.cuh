// klang-b200 — the primitive operators of the hot path over the POD state of kb_state.h.
//
// Every function is __host__ __device__: the HOST executes the event-rate halves (set()/on()/off()/
// prepare(), which draw libc rand() and call the host libm exactly like the reference does), the DEVICE
// executes the per-sample halves (process()).  There is no host implementation of a block loop anywhere
// in the product.  Built with -fmad=false -ftz=false -prec-div=true -prec-sqrt=true so fp32 arithmetic is
// IEEE and un-contracted like the reference's g++ -O3 -ffp-contract=off build.
// Citations: klang.h = nashaudio/klang v0.7.8.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "kb_math.cuh"
#include "kb_state.h"
#include "kb_rand.h"

#ifdef __CUDA_ARCH__
#define KB_SINF(x) kb_sinf(x)
#define KB_COSF(x) kb_cosf(x)
#else
#define KB_SINF(x) ::sinf(x)
#define KB_COSF(x) ::cosf(x)
#endif

#define KB_PI_F 3.14159274101257324f         /* pi.f            klang.h:227 */
#define KB_PI_INV_F 0.318309873342514038f    /* pi.inv          klang.h:97  */
#define KB_TWO_PI_F 6.28318548202514648f
#define KB_ROOT2_F 1.41421353816986084f      /* root2.f         klang.h:233 */
#define KB_ROOT2_INV_F 0.707106769084930420f /* root2.inv       klang.h:233 */
#define KB_DENORMALISE 1.175494e-38f         /* DENORMALISE     klang.h:90  */

KB_HD float kb_bits(uint32_t u) {
#ifdef __CUDA_ARCH__
	return __uint_as_float(u);
#else
	float f; memcpy(&f, &u, 4); return f;
#endif
}

// SampleRate::SampleRate                                                   klang.h:1601
KB_HD KbFs kb_make_fs(float sr) {
	KbFs fs; fs.f = sr; fs.i = (int)(sr + 0.001f); fs.inv = 1.f / sr; fs.w = 2.0f * KB_PI_F * fs.inv; fs.nyquist = sr / 2.f;
	return fs;
}

// Control::set / Control::smooth                                            klang.h:1715-1728
KB_HD float kb_clampf(float x, float lo, float hi) { return (x < lo) ? lo : (hi < x) ? hi : x; }
KB_HD void kb_control_set(KbControl& c, float x) { c.value = kb_clampf(x, c.min, c.max); }
KB_HD float kb_control_smooth(KbControl& c) { c.smoothed = c.smoothed * 0.999f + (1.f - 0.999f) * c.value; return c.smoothed; }

// ------------------------------------------------------------ Generators::Fast (klang.h:4955-5367)
// x86-64 converts float -> unsigned through a 64-bit signed conversion (wraps mod 2^32)
KB_HD uint32_t kb_f2u(float x) { return (uint32_t)(long long)x; }
// Fast::Increment::set                                                      klang.h:4968-4973
KB_HD int kb_increment_set(const KbFs& fs, float f) {
	const float FC4 = (float)261.62556530059862;
	const float FC4_FINTMAX = (float)(261.62556530059862 * 2147483648.0);
	const float FBASE = FC4_FINTMAX / fs.f;
	return (int)(2u * (uint32_t)(int)(FBASE / FC4 * f));
}
KB_HD float kb_increment_float(int amount) { return kb_bits((uint32_t)((amount >> 9) | 0x3f800000)) - 1.f; }   // klang.h:4976-4979
KB_HD uint32_t kb_phase_from_radians(float phase) { return kb_f2u(phase * 2147483648.0f / (2.f * KB_PI_F)); }  // klang.h:4993-4997 (Q2)
KB_HD float kb_phase_float(uint32_t position) { return kb_bits((position >> 9) | 0x3f800000) - 1.f; }          // klang.h:5004-5006

// OSM::init                                                                 klang.h:5206-5215
KB_HD void kb_osm_init(KbOsm& o) {
	o.state = ((o.offset - (uint32_t)o.increment) < o.duty) ? 3 : 0;
	o.f = o.delta;
	o.omf = 1.f - o.f;
	o.rcpf = 1.f / o.f;
	o.rcpf2 = 2.f * o.rcpf;
	o.col = kb_phase_float(o.duty);
	o.c1 = 1.f / o.col;
	o.c2 = -1.f / (1.0f - o.col);
}
KB_HD void kb_osm_set_duty(KbOsm& o, float duty) { o.duty = kb_phase_from_radians(duty * (2.f * KB_PI_F)); kb_osm_init(o); }  // klang.h:5246-5249
KB_HD void kb_osm_construct(KbOsm& o, int waveform, float duty) { memset(&o, 0, sizeof(o)); o.waveform = waveform; kb_osm_set_duty(o, duty); }
KB_HD void kb_osm_update_f(const KbFs& fs, KbOsm& o, float frequency) {
	o.frequency = frequency;
	o.increment = kb_increment_set(fs, frequency);
	o.delta = kb_increment_float(o.increment);
}
// OSM::set(f) / set(f,phase) / set(f,phase,duty)                            klang.h:5217-5244
KB_HD void kb_osm_set_f(const KbFs& fs, KbOsm& o, float frequency) { if (o.frequency != frequency) { kb_osm_update_f(fs, o, frequency); kb_osm_init(o); } }
KB_HD void kb_osm_set_fp(const KbFs& fs, KbOsm& o, float frequency, float phase) {
	if (o.frequency != frequency) kb_osm_update_f(fs, o, frequency);
	o.offset = kb_phase_from_radians(phase);
	kb_osm_init(o);
}
KB_HD void kb_osm_set_fpd(const KbFs& fs, KbOsm& o, float frequency, float phase, float duty) {
	if (o.frequency != frequency) kb_osm_update_f(fs, o, frequency);
	o.offset = kb_phase_from_radians(phase);
	kb_osm_set_duty(o, duty);
}
KB_HD float kb_sqr(float x) { return x * x; }

// The sample an OSM produces for transition `tr` once its phase has advanced to `offset`
// (OSM::saw / OSM::pulse, klang.h:5290-5316; g++ evaluates tick() before `offset - col`).
KB_HD float kb_osm_wave(const KbOsm& o, int tr, uint32_t offset_after) {
	const float f = o.f, omf = o.omf, rcpf = o.rcpf, rcpf2 = o.rcpf2, col = o.col, c1 = o.c1, c2 = o.c2;
	if (o.waveform == 0) {
		const float p = kb_phase_float(offset_after) - col;
		switch (tr) {
		case 3: return c1 * (p + p - f) + 1.f;
		case 0: return c2 * (p + p - f) + 1.f;
		case 2: return rcpf * (c2 * kb_sqr(p) - c1 * kb_sqr(p - f)) + 1.f;
		case 5: return -rcpf * (1.f + c2 * kb_sqr(p + omf) - c1 * kb_sqr(p)) + 1.f;
		case 7: return -rcpf * (1.f + c1 * omf * (p + p + omf)) + 1.f;
		case 4: return -rcpf * (1.f + c2 * omf * (p + p + omf)) + 1.f;
		default: return 0.f;
		}
	} else {
		const float p = kb_phase_float(offset_after);
		switch (tr) {
		case 3: return 1.f;
		case 0: return -1.f;
		case 2: return rcpf2 * (col - p) + 1.f;
		case 5: return rcpf2 * p - 1.f;
		case 7: return rcpf2 * (col - 1.0f) + 1.f;
		case 4: return rcpf2 * col - 1.f;
		default: return 0.f;
		}
	}
}
// OSM::tick + saw()/pulse(): the sequential form                            klang.h:5251-5263
KB_HD float kb_osm_tick(KbOsm& o) {
	o.state = ((o.state << 1) | (o.offset < o.duty ? 1 : 0)) & 3;
	const int tr = o.state | (o.offset < (uint32_t)o.increment ? 4 : 0);
	o.offset += (uint32_t)o.increment;
	return kb_osm_wave(o, tr, o.offset);
}
// The same sample in closed form: the phase is an integer ramp, so tick number i (0-based, counted from
// the state `o` describes) depends only on offset+i*inc and on the compare bit of the tick before it.
// Bit-identical to calling kb_osm_tick i+1 times; lets kernels evaluate oscillators in parallel over time.
KB_HD float kb_osm_at(const KbOsm& o, uint32_t i) {
	const uint32_t inc = (uint32_t)o.increment;
	const uint32_t off = o.offset + i * inc;
	if (o.waveform == 0 && o.duty == 0u && (o.state & 1) == 0) {
		// plain Saw (duty 0): `offset < duty` is never true, so the state machine only ever sees Down (0) and, on the
		// tick where the phase wraps, DownUpDown (4) — the same two expressions the generic path would select
		const float p = kb_phase_float(off + inc) - o.col;
		return (off < inc) ? -o.rcpf * (1.f + o.c2 * o.omf * (p + p + o.omf)) + 1.f : o.c2 * (p + p - o.f) + 1.f;
	}
	const int prev = (i == 0) ? (o.state & 1) : ((off - inc) < o.duty ? 1 : 0);
	const int st = (prev << 1) | (off < o.duty ? 1 : 0);
	const int tr = st | (off < inc ? 4 : 0);
	return kb_osm_wave(o, tr, off + inc);
}
// state after n ticks
KB_HD void kb_osm_advance(KbOsm& o, uint32_t n) {
	if (n == 0) return;
	const uint32_t inc = (uint32_t)o.increment;
	const uint32_t last = o.offset + (n - 1) * inc;
	const int prev = (n == 1) ? (o.state & 1) : ((last - inc) < o.duty ? 1 : 0);
	o.state = (prev << 1) | (last < o.duty ? 1 : 0);
	o.offset = last + inc;
}

// ------------------------------------------- Generic::Oscillator + Generators::Basic (klang.h:2849-2880, 4899-4944)
// Generators::Fast::Sine (klang.h:5135-5172): uint32 phase, fastsinp (5117-5132), polysin (5093-5096); Q3: set(f) is a no-op
// while f equals the cached frequency, which starts at 1000 with a zero increment (klang.h:2855)
KB_HD void kb_fsine_init(KbFastSine& o) { o.frequency = 1000.f; o.increment = 0; o.position = 0u; o.offset = 0u; }
KB_HD void kb_fsine_set_f(const KbFs& fs, KbFastSine& o, float f) { if (f != o.frequency) { o.frequency = f; o.increment = kb_increment_set(fs, f); } }
KB_HD void kb_fsine_set_fp(const KbFs& fs, KbFastSine& o, float f, float phase) {
	o.position = kb_phase_from_radians(phase); o.offset = kb_phase_from_radians(0.f); kb_fsine_set_f(fs, o, f);
}
KB_HD float kb_fsine_value(uint32_t phase) {                                                       // fastsinp(position + offset)
	float x = (kb_bits((phase >> 9) | 0x3f800000) - 1.f) * KB_TWO_PI_F;                               // fast_modp  klang.h:1424-1428
	if (x > 3.f / 2.f * KB_PI_F) x -= KB_TWO_PI_F; else if (x > KB_PI_F / 2.f) x = KB_PI_F - x;
	const float x2 = x * x;
	return (((-0.00018542f * x2 + 0.0083143f) * x2 - 0.16666f) * x2 + 1.0f) * x;
}
KB_HD float kb_fsine_tick(KbFastSine& o) {
	const float out = kb_fsine_value(o.position + o.offset);
	o.position += (uint32_t)o.increment;
	return out;
}
KB_HD void kb_bosc_init(KbBasicOsc& o) { o.increment = 0.f; o.position = 0.f; o.frequency = 1000.f; o.offset = 0.f; o.duty = 0.5f; }
KB_HD void kb_bosc_set_f(const KbFs& fs, KbBasicOsc& o, float f) { o.frequency = f; o.increment = f * 2.f * KB_PI_F / fs.f; }
KB_HD void kb_bosc_set_fp(const KbFs& fs, KbBasicOsc& o, float f, float phase) { o.position = phase; kb_bosc_set_f(fs, o, f); }
KB_HD void kb_bosc_advance(KbBasicOsc& o) {               // Phase::operator+=(float)  klang.h:1518-1525
	if (o.increment >= (2 * KB_PI_F)) return;
	o.position += o.increment;
	if (o.position > (2 * KB_PI_F)) o.position -= (2 * KB_PI_F);
}
KB_HD float kb_bosc_sine_tick(KbBasicOsc& o) {            // Basic::Sine::process (sin binds to sinf, Q10)  klang.h:4899-4903
	const float out = KB_SINF(o.position + o.offset);
	kb_bosc_advance(o);
	return out;
}
// Basic::{Saw,Triangle,Square,Pulse}::process (aliased shapes; abs == fabsf, Q4)     klang.h:4907-4944
enum { KB_BOSC_SAW = 1, KB_BOSC_TRIANGLE, KB_BOSC_SQUARE, KB_BOSC_PULSE };
KB_HD float kb_bosc_shape_tick(KbBasicOsc& o, int shape) {
	float out;
	switch (shape) {
	case KB_BOSC_SAW: out = o.position * KB_PI_INV_F - 1.f; break;
	case KB_BOSC_TRIANGLE: out = fabsf(2.f * o.position * KB_PI_INV_F - 2) - 1.f; break;
	case KB_BOSC_SQUARE: out = o.position > KB_PI_F ? 1.f : -1.f; break;
	default: out = o.position > (o.duty * KB_PI_F) ? 1.f : -1.f; break;
	}
	kb_bosc_advance(o);
	return out;
}

// ------------------------------------------------------------ Filters (klang.h:5383-5813)
enum { KB_BQ_LPF = 0, KB_BQ_HPF, KB_BQ_BPF, KB_BQ_BRF, KB_BQ_APF, KB_BQ_BW2 };
KB_HD void kb_biquad_construct(KbBiquad& b, int type) { memset(&b, 0, sizeof(b)); b.type = type; b.b0 = 1.f; b.cos0 = 1.f; }
KB_HD void kb_biquad_reset(KbBiquad& b) { b.f = 0; b.Q = 0; b.b0 = 1; b.a1 = b.a2 = b.b1 = b.b2 = 0; b.a = 0; b.z0 = b.z1 = 0; }   // klang.h:5565-5572
// constant{x}.inv = float(1.0 / double(x))                                  klang.h:96-98
KB_HD float kb_const_inv(float x) {
#ifdef __CUDA_ARCH__
	// (float)(1.0 / (double)x) equals the correctly rounded fp32 reciprocal: rounding the quotient of 24-bit operands through a
	// 53-bit intermediate is innocuous (53 >= 2*24 + 2; 0 mismatches over the 2^24 floats of [0.5, 2) on the host)
	const float ax = fabsf(x);
	if (ax >= 1e-30f && ax <= 1e30f) return __frcp_rn(x);
#endif
	const double v = (double)x; return v == 0.0 ? 0.0f : (float)(1.0 / v);
}
// LPF::init 5658-5665, HPF::init 5675-5682, BPF 5720-5729, BRF 5734-5739, Butterworth::LPF<2> 5803-5810
KB_HD void kb_biquad_init(KbBiquad& b) {
	const float inv = kb_const_inv(1.f + b.a);
	const float cos0 = b.cos0, a = b.a;
	switch (b.type) {
	case KB_BQ_LPF:
		b.a1 = inv * (-2.f * cos0); b.a2 = inv * (1.f - a);
		b.b2 = b.b0 = inv * (1.f - cos0) * 0.5f; b.b1 = inv * (1.f - cos0); break;
	case KB_BQ_HPF:
		b.a1 = inv * (-2.f * cos0); b.a2 = inv * (1.f - a);
		b.b2 = b.b0 = inv * (1.f + cos0) * 0.5f; b.b1 = inv * -(1.f + cos0); break;
	case KB_BQ_BPF:
		b.a1 = inv * (-2.f * cos0); b.a2 = inv * (1.f - a);
		b.b0 = inv * a; b.b1 = 0; b.b2 = inv * -a; break;
	case KB_BQ_BRF:
		b.b1 = b.a1 = inv * (-2.f * cos0); b.a2 = inv * (1.f - a); b.b0 = b.b2 = inv; break;
	case KB_BQ_BW2:
		b.b0 = inv * ((1.f - cos0) / 2.f); b.b1 = inv * (1.f - cos0); b.b2 = inv * ((1.f - cos0) / 2.f);
		b.a1 = inv * (-2.f * cos0); b.a2 = inv * (1.f - a); break;
	default: break;
	}
}
// Filter::set(f,Q)                                                          klang.h:5584-5600
KB_HD void kb_biquad_set(const KbFs& fs, KbBiquad& b, float f, float Q) {
	if (Q < 0) Q = f / -Q;
	if (b.f != f || b.Q != Q) {
		b.f = f; b.Q = Q;
		const float w = f * fs.w;
		b.cos0 = KB_COSF(w);
		b.sin0 = KB_SINF(w);
		if (Q < 0.5) Q = 0.5f;
		b.a = b.sin0 / (2.f * Q);
		kb_biquad_init(b);
	}
}
// APF::set(f, r) + APF::init                                                klang.h:5752-5772
KB_HD void kb_apf_set(const KbFs& fs, KbBiquad& b, float f, float r) {
	if (b.f != f || b.a != r) {
		b.f = f; b.a = r;
		const float w = f * fs.w;
		b.cos0 = KB_COSF(w); b.sin0 = KB_SINF(w);
		const float omega = 2.0f * KB_PI_F * f / fs.f;
		const float c0 = KB_COSF(omega);
		b.b0 = b.a2 = b.a * b.a;
		b.b1 = b.a1 = (-2.f * b.a * c0);
		b.b2 = 1.f;
	}
}
KB_HD void kb_biquad_set_f(const KbFs& fs, KbBiquad& b, float f) {                                               // klang.h:5575
	if (b.type == KB_BQ_APF) kb_apf_set(fs, b, f, 1.f); else kb_biquad_set(fs, b, f, KB_ROOT2_INV_F);
}
// Filter::process                                                           klang.h:5605-5612
KB_HD float kb_biquad_tick(KbBiquad& b, float in) {
	const float z0 = b.z0, z1 = b.z1;
	const float y = b.b0 * in + z0;
	b.z0 = b.b1 * in - b.a1 * y + z1;
	b.z1 = b.b2 * in - b.a2 * y;
	return y;
}

enum { KB_OP_LPF = 0, KB_OP_HPF, KB_OP_BW1 };
KB_HD void kb_onepole_construct(KbOnePole& p, int type) { memset(&p, 0, sizeof(p)); p.type = type; p.b0 = 1.f; }
KB_HD void kb_onepole_reset(KbOnePole& p) { p.a1 = 0; p.b0 = 1; p.b1 = 0; p.f = 0; p.z = 0; }   // klang.h:5482-5488
// OnePole::LPF::init / HPF::init — control-rate only (expf from the host libm)    klang.h:5508-5512, 5535-5541
// (on the device — translated programs that set the filter per sample — expf is kb_expf, the restatement of the host's; tanf has none)
KB_HD void kb_onepole_set(const KbFs& fs, KbOnePole& p, float f) {
	if (p.f != f) {
		p.f = f;
		if (p.type == KB_OP_BW1) {                                             // Butterworth::LPF<1>::init  klang.h:5786-5793
			const float c = 1.f / ::tanf(KB_PI_F * f * fs.inv);
			const float inv = kb_const_inv(1.f + c);
			p.b0 = inv; p.a1 = (1.f - c) * inv;
			return;
		}
#ifdef __CUDA_ARCH__
		const float e = kb_expf(-f * fs.w);
#else
		const float e = ::expf(-f * fs.w);
#endif
		if (p.type == KB_OP_LPF) { p.b0 = 1 - e; p.a1 = e; }
		else { p.b0 = 0.5f * (1.f + e); p.b1 = -p.b0; p.a1 = e; }
	}
}
KB_HD float kb_onepole_tick(KbOnePole& p, float in) {
	if (p.type == KB_OP_LPF) p.out = p.b0 * in + p.a1 * p.out + KB_DENORMALISE;                            // klang.h:5515-5517
	else if (p.type == KB_OP_HPF) { p.out = p.b0 * in + p.b1 * p.z + p.a1 * p.out + KB_DENORMALISE; p.z = in; }   // klang.h:5499-5502
	else { p.out = p.b0 * (in + p.z) - p.a1 * p.out; p.z = in; }                                           // Butterworth::LPF<1>  klang.h:5795-5798
	return p.out;
}

// ------------------------------------------------------------ Envelope / ADSR (klang.h:3723-4137)
KB_HD void kb_ramp_set_value(KbEnv& e, float v) { e.r_out = v; e.r_target = v; e.r_active = 0; }        // klang.h:3763-3767
KB_HD void kb_ramp_set_target(KbEnv& e, float t) { e.r_target = t; e.r_active = (e.r_out != t); }       // klang.h:3757-3760
// Envelope::setTargetTime (abs == fabsf, Q4)                                klang.h:4077-4081
KB_HD void kb_env_set_target(const KbFs& fs, KbEnv& e, float x, float y, float time) {
	e.time = time;
	kb_ramp_set_target(e, y);
	e.r_rate = fabsf(y - e.r_out) / ((x - time) * fs.f);
}
// Envelope::initialise                                                      klang.h:3974-3989
KB_HD void kb_env_initialise(const KbFs& fs, KbEnv& e) {
	e.point = 0;
	e.timeInc = 1.0f / fs.f;
	e.loop_start = e.loop_end = -1;
	e.stage = KB_ENV_SUSTAIN;
	if (e.npoints) {
		e.out = e.py[0];
		kb_ramp_set_value(e, e.py[0]);
		if (e.npoints > 1) kb_env_set_target(fs, e, e.px[1], e.py[1], e.px[0]);
	} else {
		e.out = 1.0f;
		kb_ramp_set_value(e, 1.0f);
	}
}
KB_HD void kb_env_construct(const KbFs& fs, KbEnv& e) {          // Envelope::Envelope(): one point (0,1)   klang.h:3867
	memset(&e, 0, sizeof(e));
	e.npoints = 1; e.px[0] = 0.f; e.py[0] = 1.f;
	kb_env_initialise(fs, e);
}
KB_HD void kb_env_set_points(const KbFs& fs, KbEnv& e, int n, const float* xy) {   // klang.h:3887-3896
	e.npoints = n;
	for (int p = 0; p < n; p++) { e.px[p] = xy[2 * p]; e.py[p] = xy[2 * p + 1]; }
	kb_env_initialise(fs, e);
}
KB_HD void kb_env_set_loop(KbEnv& e, int s, int t) { if (s >= 0 && t < e.npoints) { e.loop_start = s; e.loop_end = t; } }   // klang.h:3923-3926
KB_HD void kb_env_release(const KbFs& fs, KbEnv& e, float time, float level) { e.stage = KB_ENV_RELEASE; kb_env_set_target(fs, e, time, level, 0.f); }   // klang.h:3961-3966
// Envelope::process with Linear::operator++                                 klang.h:4018-4051, 3785-3806
KB_HD float kb_env_tick(const KbFs& fs, KbEnv& e) {
	const float output = e.r_out;
	if (e.r_active) {
		if (e.r_target > e.r_out) {
			e.r_out += e.r_rate;
			if (e.r_out >= e.r_target) { e.r_out = e.r_target; e.r_active = 0; }
		} else {
			e.r_out -= e.r_rate;
			if (e.r_out <= e.r_target) { e.r_out = e.r_target; e.r_active = 0; }
		}
	}
	e.out = output;
	if (e.stage == KB_ENV_SUSTAIN) {
		e.time += e.timeInc;
		if (!e.r_active) {
			const bool loop_active = e.loop_start != -1 && e.loop_end != -1;
			if (loop_active && (e.point + 1) >= e.loop_end) {
				e.point = e.loop_start;
				kb_ramp_set_value(e, e.py[e.point]);
				if (e.loop_start != e.loop_end)
					kb_env_set_target(fs, e, e.px[e.point + 1], e.py[e.point + 1], e.px[e.point]);
			} else if ((e.point + 1) < e.npoints) {
				if (e.time >= e.px[e.point + 1]) {
					e.point++;
					kb_ramp_set_value(e, e.py[e.point]);
					if ((e.point + 1) < e.npoints)
						kb_env_set_target(fs, e, e.px[e.point + 1], e.py[e.point + 1], e.px[e.point]);
				}
			} else {
				e.stage = KB_ENV_OFF;
			}
		}
	} else if (e.stage == KB_ENV_RELEASE) {
		if (!e.r_active) e.stage = KB_ENV_OFF;
	}
	return e.out;
}
// The same Envelope::process over a register-resident scalar state with the breakpoints held elsewhere (shared
// memory): the breakpoint arrays are only touched when a ramp segment ends, so the per-sample path is a handful of
// register operations.  Bit-identical to kb_env_tick.
struct KbEnvR { float r_out, r_target, r_rate, time, timeInc, out; int r_active, stage, point, loop_start, loop_end, npoints; };
KB_HD void kb_envr_load(KbEnvR& r, const KbEnv& e) {
	r.r_out = e.r_out; r.r_target = e.r_target; r.r_rate = e.r_rate; r.time = e.time; r.timeInc = e.timeInc; r.out = e.out;
	r.r_active = e.r_active; r.stage = e.stage; r.point = e.point; r.loop_start = e.loop_start; r.loop_end = e.loop_end; r.npoints = e.npoints;
}
KB_HD void kb_envr_store(const KbEnvR& r, KbEnv& e) {
	e.r_out = r.r_out; e.r_target = r.r_target; e.r_rate = r.r_rate; e.time = r.time; e.out = r.out;
	e.r_active = r.r_active; e.stage = r.stage; e.point = r.point;
}
KB_HD void kbr_set_value(KbEnvR& e, float v) { e.r_out = v; e.r_target = v; e.r_active = 0; }
KB_HD void kbr_set_target(const KbFs& fs, KbEnvR& e, float x, float y, float time) {
	e.time = time;
	e.r_target = y; e.r_active = (e.r_out != y);
	e.r_rate = fabsf(y - e.r_out) / ((x - time) * fs.f);
}
// the part of Envelope::process that follows the ramp step (stage logic)      klang.h:4024-4050
KB_HD void kbr_after_ramp(const KbFs& fs, KbEnvR& e, const float* px, const float* py) {
	if (e.stage == KB_ENV_SUSTAIN) {
		e.time += e.timeInc;
		if (!e.r_active) {
			const bool loop_active = e.loop_start != -1 && e.loop_end != -1;
			if (loop_active && (e.point + 1) >= e.loop_end) {
				e.point = e.loop_start;
				kbr_set_value(e, py[e.point]);
				if (e.loop_start != e.loop_end) kbr_set_target(fs, e, px[e.point + 1], py[e.point + 1], px[e.point]);
			} else if ((e.point + 1) < e.npoints) {
				if (e.time >= px[e.point + 1]) {
					e.point++;
					kbr_set_value(e, py[e.point]);
					if ((e.point + 1) < e.npoints) kbr_set_target(fs, e, px[e.point + 1], py[e.point + 1], px[e.point]);
				}
			} else {
				e.stage = KB_ENV_OFF;
			}
		}
	} else if (e.stage == KB_ENV_RELEASE) {
		if (!e.r_active) e.stage = KB_ENV_OFF;
	}
}
KB_HD float kb_envr_tick(const KbFs& fs, KbEnvR& e, const float* px, const float* py) {
	const float output = e.r_out;
	if (e.r_active) {
		if (e.r_target > e.r_out) {
			e.r_out += e.r_rate;
			if (e.r_out >= e.r_target) { e.r_out = e.r_target; e.r_active = 0; }
		} else {
			e.r_out -= e.r_rate;
			if (e.r_out <= e.r_target) { e.r_out = e.r_target; e.r_active = 0; }
		}
	}
	e.out = output;
	kbr_after_ramp(fs, e, px, py);
	return output;
}
// `steps` consecutive ticks into row[0..steps).  The envelope spends almost all of its time in one of four steady modes —
// a running ramp (up or down), the ADSR sustain hold (a one-point loop re-asserting its level every sample), waiting for
// the next breakpoint's time, or Off.  All four are the SAME four-ticks-per-iteration loop with per-lane constants
// (signed rate or -0, time increment or -0, and the exit test), so the lanes of a warp — voices in different envelope
// phases — run it together instead of one mode after the other.  x + (-0) == x bit for bit, which is how a mode
// switches an update off.  Every mode change (ramp crossing, breakpoint reached, loop jump, release end) and every tail
// shorter than four ticks goes through the generic tick, so the sequence of values is bit-identical to kb_envr_tick.
KB_HD uint32_t kb_fbits(float f) {
#ifdef __CUDA_ARCH__
	return __float_as_uint(f);
#else
	uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
KB_HD void kb_envr_run(const KbFs& fs, KbEnvR& e, const float* px, const float* py, float* row, int steps) {
	int t = 0;
	while (t < steps) {
		if (t + 4 <= steps) {
			const bool act = e.r_active != 0, sus = e.stage == KB_ENV_SUSTAIN;
			const bool at_loop_end = e.loop_start != -1 && e.loop_end != -1 && (e.point + 1) >= e.loop_end;
			// ramp: a positive finite rate moves r_out monotonically towards the target, so the direction test of
			// Linear::operator++ (klang.h:3785-3806) is invariant until the ramp crosses
			const bool is_ramp = act && e.stage != KB_ENV_OFF && e.r_rate > 0.f && e.r_rate <= 3.0e38f;
			const bool is_off = !act && e.stage == KB_ENV_OFF;
			const bool is_wait = !act && sus && !at_loop_end && (e.point + 1) < e.npoints;       // klang.h:4036-4043
			bool is_hold = false;                                                                // klang.h:4029-4035 at its fixed point
			if (!act && sus && at_loop_end && e.loop_start == e.loop_end && e.point == e.loop_start) {
				const uint32_t lvl = kb_fbits(py[e.loop_start]);
				is_hold = lvl == kb_fbits(e.r_out) && lvl == kb_fbits(e.r_target);
			}
			if (is_ramp || is_off || is_wait || is_hold) {
				const bool up = e.r_target > e.r_out;
				const float srate = is_ramp ? (up ? e.r_rate : -e.r_rate) : -0.f;
				const float tinc = sus ? e.timeInc : -0.f;
				// the exit test as three comparisons against per-lane constants (no branches between the modes): a rising ramp stays
				// while r4 < target, a falling one while r4 > target, a wait while t4 < x; everything else compares against +-inf.
				// A NaN or infinite r4 / t4 merely sends the lane through the generic tick.
				const float inf = kb_bits(0x7f800000u);
				const float r_hi = (is_ramp && up) ? e.r_target : inf, r_lo = (is_ramp && !up) ? e.r_target : -inf;
				const float t_hi = is_wait ? px[e.point + 1] : inf;
				float r = e.r_out, time = e.time, last = e.out;
				// the partial sums are formed exactly as single ticks would form them; r and time only move one way, so "the last
				// of the group has not crossed / arrived" implies none has.  Groups of 32 first (the exit test and its branch cost
				// as much as a dozen dependent adds), then groups of 16 and of 4 up to the mode change.
				while (t + 32 <= steps) {
					float rr[33], tt = time;
					rr[0] = r;
					#pragma unroll
					for (int j = 0; j < 32; j++) { rr[j + 1] = rr[j] + srate; tt = tt + tinc; }
					const bool stay = (rr[32] < r_hi) & (rr[32] > r_lo) & (tt < t_hi);
					if (!stay) break;
					#pragma unroll
					for (int j = 0; j < 32; j++) row[t + j] = rr[j];
					last = rr[31]; r = rr[32]; time = tt; t += 32;
				}
				while (t + 16 <= steps) {
					float rr[17], tt = time;
					rr[0] = r;
					#pragma unroll
					for (int j = 0; j < 16; j++) { rr[j + 1] = rr[j] + srate; tt = tt + tinc; }
					const bool stay = (rr[16] < r_hi) & (rr[16] > r_lo) & (tt < t_hi);
					if (!stay) break;
					#pragma unroll
					for (int j = 0; j < 16; j++) row[t + j] = rr[j];
					last = rr[15]; r = rr[16]; time = tt; t += 16;
				}
				while (t + 4 <= steps) {
					const float r1 = r + srate, r2 = r1 + srate, r3 = r2 + srate, r4 = r3 + srate;
					const float t1 = time + tinc, t2 = t1 + tinc, t3 = t2 + tinc, t4 = t3 + tinc;
					const bool stay = (r4 < r_hi) & (r4 > r_lo) & (t4 < t_hi);
					if (!stay) break;
					row[t] = r; row[t + 1] = r1; row[t + 2] = r2; row[t + 3] = r3;
					last = r3; r = r4; time = t4; t += 4;
				}
				e.r_out = r; e.time = time; e.out = last;
				if (t >= steps) break;
			}
		}
		row[t++] = kb_envr_tick(fs, e, px, py);
	}
}
// ---- the same run in UNIFORM GROUPS of 16 ticks (round 2, kb_sub_flow_kernel).  kb_envr_run leaves its fast loop through ever smaller
// groups and single ticks at a mode change, which costs the lane that changes ~5000 cycles while the other lanes of the warp wait
// (measured: profiles/r02_c2_trace.txt, ticks 3 / 11 / 29).  Here every lane walks the same loop: per group a lane either takes the 16
// branch-free steps of its steady mode or — when it is in no steady mode, or the group's last value crosses the mode's bound — 16
// generic ticks, after which its mode constants are re-derived; the warp reconverges after every group, so a mode change costs one
// group of generic ticks.  The values are those of kb_envr_tick bit for bit (same argument as kb_envr_run: r and time move one way,
// so "the value after the group has not crossed / arrived" implies no tick of the group changed mode).
struct KbEnvMode { float srate, tinc, r_hi, r_lo, t_hi; bool fast; };
KB_HD KbEnvMode kb_envr_mode(const KbEnvR& e, const float* px, const float* py) {
	const bool act = e.r_active != 0, sus = e.stage == KB_ENV_SUSTAIN;
	const bool at_loop_end = e.loop_start != -1 && e.loop_end != -1 && (e.point + 1) >= e.loop_end;
	const bool is_ramp = act && e.stage != KB_ENV_OFF && e.r_rate > 0.f && e.r_rate <= 3.0e38f;
	const bool is_off = !act && e.stage == KB_ENV_OFF;
	const bool is_wait = !act && sus && !at_loop_end && (e.point + 1) < e.npoints;
	bool is_hold = false;
	if (!act && sus && at_loop_end && e.loop_start == e.loop_end && e.point == e.loop_start) {
		const uint32_t lvl = kb_fbits(py[e.loop_start]);
		is_hold = lvl == kb_fbits(e.r_out) && lvl == kb_fbits(e.r_target);
	}
	const bool up = e.r_target > e.r_out;
	const float inf = kb_bits(0x7f800000u);
	KbEnvMode m;
	m.fast = is_ramp || is_off || is_wait || is_hold;
	m.srate = is_ramp ? (up ? e.r_rate : -e.r_rate) : -0.f;
	m.tinc = sus ? e.timeInc : -0.f;
	m.r_hi = (is_ramp && up) ? e.r_target : inf;
	m.r_lo = (is_ramp && !up) ? e.r_target : -inf;
	m.t_hi = is_wait ? px[e.point + 1] : inf;
	return m;
}
// ALIGNED16: `row` is 16-byte aligned (the groups are then stored as four 128-bit words on the device)
template <bool ALIGNED16>
KB_HD void kb_envr_run16(const KbFs& fs, KbEnvR& e, const float* px, const float* py, float* row, int steps) {
	int t = 0;
	KbEnvMode m = kb_envr_mode(e, px, py);
	for (; t + 16 <= steps; t += 16) {
		bool done = false;
		if (m.fast) {
			float rr[17], tt = e.time;
			rr[0] = e.r_out;
			#pragma unroll
			for (int j = 0; j < 16; j++) { rr[j + 1] = rr[j] + m.srate; tt = tt + m.tinc; }
			if ((rr[16] < m.r_hi) & (rr[16] > m.r_lo) & (tt < m.t_hi)) {
#ifdef __CUDA_ARCH__
				if (ALIGNED16) {
					#pragma unroll
					for (int q = 0; q < 4; q++) *reinterpret_cast<float4*>(row + t + 4 * q) = make_float4(rr[4 * q], rr[4 * q + 1], rr[4 * q + 2], rr[4 * q + 3]);
				} else
#endif
				{
					#pragma unroll
					for (int j = 0; j < 16; j++) row[t + j] = rr[j];
				}
				e.out = rr[15]; e.r_out = rr[16]; e.time = tt;
				done = true;
			}
		}
		if (!done) {
			for (int j = 0; j < 16; j++) row[t + j] = kb_envr_tick(fs, e, px, py);
			m = kb_envr_mode(e, px, py);
		}
	}
	for (; t < steps; t++) row[t] = kb_envr_tick(fs, e, px, py);
}
// ---- a whole tile at once.  Even the uniform groups pay ~10 cycles per tick, most of it the exit test and its branch after every group
// (tools/micro/env_floor.cu: the two FADD chains and the stores alone cost 6.4).  r and time move one way, so ONE test after the tile's
// last tick decides for all of them: a lane in a steady mode runs the tile's ticks straight through, storing as it goes, and keeps the
// result if the value after the last tick has not crossed the mode's bound; otherwise — a few tiles per note — the tile is done again
// from the untouched state by kb_envr_run16, whose generic ticks overwrite the row.  `steps` must be a multiple of 16 for the straight
// run (else kb_envr_run16 does the tile).  Bit-identical to kb_envr_tick.
template <bool ALIGNED16>
KB_HD void kb_envr_run_tile(const KbFs& fs, KbEnvR& e, const float* px, const float* py, float* row, int steps) {
	if (steps > 0 && (steps & 15) == 0) {
		const KbEnvMode m = kb_envr_mode(e, px, py);
		if (m.fast) {
			float r = e.r_out, tt = e.time, last = e.out;
			for (int t = 0; t < steps; t += 16) {
				float rr[17];
				rr[0] = r;
				#pragma unroll
				for (int j = 0; j < 16; j++) { rr[j + 1] = rr[j] + m.srate; tt = tt + m.tinc; }
#ifdef __CUDA_ARCH__
				if (ALIGNED16) {
					#pragma unroll
					for (int q = 0; q < 4; q++) *reinterpret_cast<float4*>(row + t + 4 * q) = make_float4(rr[4 * q], rr[4 * q + 1], rr[4 * q + 2], rr[4 * q + 3]);
				} else
#endif
				{
					#pragma unroll
					for (int j = 0; j < 16; j++) row[t + j] = rr[j];
				}
				last = rr[15]; r = rr[16];
			}
			if ((r < m.r_hi) & (r > m.r_lo) & (tt < m.t_hi)) { e.out = last; e.r_out = r; e.time = tt; return; }
		}
	}
	kb_envr_run16<ALIGNED16>(fs, e, px, py, row, steps);
}
// Envelope::at                                                              klang.h:3929-3942
KB_HD float kb_env_at(const float* px, const float* py, int npoints, float time) {
	if (npoints == 0) return 0;
	float lx = 0, ly = py[0];
	for (int p = 0; p < npoints; p++) {
		if (px[p] >= time) {
			const float dx = px[p] - lx;
			const float dy = py[p] - ly;
			const float x = time - lx;
			return dx == 0 ? ly : (ly + x * dy / dx);
		}
		lx = px[p]; ly = py[p];
	}
	return py[npoints - 1];
}
// ADSR::set / ADSR::ADSR / ADSR::release                                    klang.h:4113-4132
KB_HD void kb_adsr_set(const KbFs& fs, KbEnv& e, float attack, float decay, float sustain, float release) {
	e.A = attack; e.D = decay + 0.005f; e.S = sustain; e.R = release + 0.005f;
	e.npoints = 3;
	e.px[0] = 0; e.py[0] = 0;
	e.px[1] = e.A; e.py[1] = 1;
	e.px[2] = e.A + e.D; e.py[2] = e.S;
	kb_env_initialise(fs, e);
	kb_env_set_loop(e, 2, 2);
}
KB_HD void kb_adsr_construct(const KbFs& fs, KbEnv& e) { kb_env_construct(fs, e); kb_adsr_set(fs, e, 0.5f, 0.5f, 1.f, 0.5f); }
KB_HD void kb_adsr_release(const KbFs& fs, KbEnv& e) { kb_env_release(fs, e, e.R, 0.f); }

// ------------------------------------------------------------ Envelope::Follower::Window<64> (klang.h:5904-5948)
// mean (rms = 0) / rms over a 64-sample moving sum kept in a DOUBLE (klang.h:5909), `sum * window.inv` rounded to float, sqrt for
// rms, then AR::process (klang.h:5881-5884) with the attack / release coefficients A, R from the host (AR::set is event-rate code).
// Host + device: tests/host/window_check.cpp renders it with g++ against the golden vectors.
KB_HD void kb_window_follower_run(int rms, float A, float R, int n, const float* in, float* out, float* coeffs) {
	const float inv = (float)(1.0 / 64.0);
	float buf[64];
	for (int i = 0; i < 64; i++) buf[i] = 0.f;
	int pos = 0;
	double sum = 0;
	float ar = 0.f;
	for (int s = 0; s < n; s++) {
		sum -= (double)buf[pos];
		buf[pos] = rms ? in[s] * in[s] : fabsf(in[s]);
		sum += (double)buf[pos];
		if (++pos == 64) pos = 0;
		float x = (float)(sum * inv);
		if (rms) x = sqrtf(x);
		const float smoothing = x > ar ? A : R;
		ar = ar + smoothing * (x - ar);
		out[s] = ar;
	}
	coeffs[0] = A; coeffs[1] = R; coeffs[2] = ar; coeffs[3] = (float)sum; coeffs[4] = 0.f;
}

// ------------------------------------------------------------ Delay (klang.h:3381-3512), ring in HBM
KB_HD void kb_delay_construct(KbDelay& d, int size, long long ring) {
	d.SIZE = size; d.time = 1; d.position = 0; d.last_position = 0; d.last_fraction = 0.f; d.out = 0.f; d.ring = ring;
}
KB_HD void kb_delay_set(KbDelay& d, float samples) {                         // Delay::set  klang.h:3480-3489
	d.time = samples < d.SIZE ? (float)samples : d.SIZE;
	float read = (float)(d.position - 1) - d.time;
	if (read < 0.f) read += d.SIZE;
	d.last_position = (int)read;
	d.last_fraction = read - d.last_position;
}
KB_D void kb_delay_write(KbDelay& d, float* ring, float in) {                // Delay::input  klang.h:3396-3403
	ring[d.position] = in;
	d.position++;
	if (d.position == d.SIZE) d.position = 0;
}
KB_D float kb_delay_tap_f(const KbDelay& d, const float* ring, float delay) {   // Delay::tap(float)  klang.h:3412-3427
	float read = (float)(d.position - 1) - delay;
	if (read < 0.f) read += d.SIZE;
	const int i = (int)read;
	const float fraction = read - i;
	const int j = (i + 1) % d.SIZE;
	return ring[i] + fraction * (ring[j] - ring[i]);
}
KB_D float kb_delay_tap_i(const KbDelay& d, const float* ring, int delay) {     // Delay::tap(int)  klang.h:3405-3410
	int read = (d.position - 1) - delay;
	if (read < 0) read += d.SIZE;
	return ring[read];
}
KB_D float kb_delay_lagrange(const KbDelay& d, const float* ring, float delay) {   // Delay::lagrange (third order)  klang.h:3429-3458
	float read = (float)(d.position - 1) - delay;
	if (read < 0.f) read += d.SIZE;
	const int i = (int)read;
	const float x = read - i;
	const float y0 = ring[(i - 1 + d.SIZE) % d.SIZE], y1 = ring[i], y2 = ring[(i + 1) % d.SIZE], y3 = ring[(i + 2) % d.SIZE];
	const float c0 = (-x * (x - 1) * (x - 2)) / 6.0f;
	const float c1 = ((x + 1) * (x - 1) * (x - 2)) / 2.0f;
	const float c2 = (-x * (x + 1) * (x - 2)) / 2.0f;
	const float c3 = (x * (x + 1) * (x - 1)) / 6.0f;
	return c0 * y0 + c1 * y1 + c2 * y2 + c3 * y3;
}
// The Delay<1000> known-answer loop of tests/cases.py (one object, sample by sample): write in[s]; tap(int di[s]); tap(float df[s]);
// lagrange(df[s]); set(set_at[s]) when >= 0; process() once a read head exists.  Host + device (tests/host/delay_check.cpp).
KB_D float kb_delay_tick(KbDelay& d, const float* ring);
KB_D void kb_delay_kat(int n, const float* in, const int* di, const float* df, const float* set_at, float* ring,
                       float* out_i, float* out_f, float* out_p, float* out_l) {
	KbDelay d; kb_delay_construct(d, 1000, 0);
	for (int s = 0; s <= 1000; s++) ring[s] = 0.f;
	bool have_set = false;
	for (int s = 0; s < n; s++) {
		kb_delay_write(d, ring, in[s]);
		out_i[s] = kb_delay_tap_i(d, ring, di[s]);
		out_f[s] = kb_delay_tap_f(d, ring, df[s]);
		out_l[s] = kb_delay_lagrange(d, ring, df[s]);
		if (set_at[s] >= 0.f) { kb_delay_set(d, set_at[s]); have_set = true; }
		out_p[s] = have_set ? kb_delay_tick(d, ring) : 0.f;
	}
}
KB_D float kb_delay_tick(KbDelay& d, const float* ring) {                    // Delay::process  klang.h:3461-3473
	const int i = d.last_position;
	const int j = (i + 1) % d.SIZE;
	d.out = ring[i] + d.last_fraction * (ring[j] - ring[i]);
	d.last_position = (d.last_position + 1) % d.SIZE;
	return d.out;
}
// Stereo::Delay::tap(float): both channels read at the left line's position   klang.h:4668-4681
KB_D void kb_sdelay_tap_f(const KbDelay& l, const float* ringl, const float* ringr, float delay, float& ol, float& orr) {
	float read = (float)(l.position - 1) - delay;
	if (read < 0.f) read += l.SIZE;
	const float f = floorf(read);
	delay = read - f;
	const int i = (int)read;
	const int j = (i == (l.SIZE - 1)) ? 0 : (i + 1);
	ol = ringl[i] * (1.f - delay) + ringl[j] * delay;
	orr = ringr[i] * (1.f - delay) + ringr[j] * delay;
}
