/* TEST INFRASTRUCTURE — not part of the product.  See klang_port.h.
 *
 * Plain-C restatement of klang v0.7.8 (reference: /root/reference/klang.h and
 * /root/reference/examples/NAME.k; all klang.h:NNNN citations refer to that file).
 * Build: gcc -std=gnu99 -O2 -ffp-contract=off -fPIC -shared (no -march=native, no
 * -ffast-math) so float arithmetic is evaluated exactly like the reference's
 * g++ -O3 -ffp-contract=off build: IEEE fp32, no FMA contraction, no FTZ.
 */
#include "klang_port.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ===================================================================== L0 */

static const float KP_PI_F = 3.14159274101257324f;      /* constant pi .f          klang.h:227 */
static const float KP_PI_INV_F = 0.318309873342514038f; /* pi.inv = float(1/pi.d)   klang.h:97  */
static const float KP_TWO_PI_F = 6.28318548202514648f;  /* 2.f * pi                */
static const float KP_ROOT2_F = 1.41421353816986084f;   /* root2.f                 klang.h:233 */
static const float KP_ROOT2_INV_F = 0.707106769084930420f; /* root2.inv            klang.h:233 */
#define KP_DENORMALISE 1.175494e-38f                    /* klang.h:90 */

typedef struct { float f; int i; double d; float inv, w, nyquist; } kp_samplerate;
static kp_samplerate kp_fs = { 44100.f, 44100, 44100.0, 1.f / 44100.f, 0.f, 22050.f };
static int kp_fs_init = 0;

/* SampleRate::SampleRate  klang.h:1601 */
void kp_set_fs(float sr) {
	kp_fs.f = sr;
	kp_fs.i = (int)(sr + 0.001f);
	kp_fs.d = (double)sr;
	kp_fs.inv = 1.f / sr;
	kp_fs.w = 2.0f * KP_PI_F * kp_fs.inv;
	kp_fs.nyquist = sr / 2.f;
	kp_fs_init = 1;
}
static void kp_ensure_fs(void) { if (!kp_fs_init) kp_set_fs(44100.f); }
float kp_get_fs(void) { kp_ensure_fs(); return kp_fs.f; }
void kp_srand(unsigned seed) { srand(seed); }
int kp_version(void) { return 708; }

static float kp_bits(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* random<float>(min,max)  klang.h:236-237 */
static float kp_randomf(float lo, float hi) { return rand() * ((hi - lo) / (float)RAND_MAX) + lo; }
/* random<double>(min,max)  klang.h:236-237 */
static double kp_randomd(double lo, double hi) { return rand() * ((hi - lo) / (double)RAND_MAX) + lo; }

/* power(float base, float exp), second overload  klang.h:187-218 */
static float kp_powerf(float base, float e) {
	if (base == 10.f) return (float)expf(e * (float)2.3025850929940456840179914546843642076011014886287729760333279009);
	else if (e == 0.f) return 1.f;
	else if (e == 1.f) return base;
	else if (e == 2.f) return base * base;
	else if (e == 3.f) return base * base * base;
	else if (e == 4.f) return base * base * base * base;
	else if (e == -1.f) return 1.f / base;
	else if (e == -2.f) return 1.f / (base * base);
	else if (e == -3.f) return 1.f / (base * base * base);
	else if (e == -4.f) return 1.f / (base * base * base * base);
	return powf(base, e);
}

/* Pitch::operator->  klang.h:1568-1571 */
float kp_pitch_to_frequency(float p) { kp_ensure_fs(); return 440.f * kp_powerf(2.f, (p - 69.f) / 12.f); }

/* x86-64 g++ converts float → unsigned through a 64-bit signed conversion (wraps mod 2^32) */
static uint32_t kp_f2u(float x) { return (uint32_t)(long long)x; }

/* ------------------------------------------------------------ Control (1655-1755) */
typedef struct { float min, max, value, smoothed; } kp_control;
static kp_control kp_dial(float lo, float hi, float initial) { kp_control c = { lo, hi, initial, 0.f }; return c; } /* klang.h:1796-1799 */
static float kp_clampf(float x, float lo, float hi) { return (x < lo) ? lo : (hi < x) ? hi : x; } /* std::clamp */
static void kp_control_set(kp_control* c, float x) { c->value = kp_clampf(x, c->min, c->max); }  /* klang.h:1725-1728 */
static float kp_control_smooth1(kp_control* c) {                                                  /* klang.h:1715-1716 */
	c->smoothed = c->smoothed * 0.999f + (1.f - 0.999f) * c->value;
	return c->smoothed;
}

/* ================================================= Generators::Fast (4955-5367) */

/* Fast::Increment::set  klang.h:4968-4973 */
static int kp_increment_set(float f) {
	const float FC4 = (float)261.62556530059862;
	const float FC4_FINTMAX = (float)(261.62556530059862 * 2147483648.0);
	const float FBASE = FC4_FINTMAX / kp_fs.f;
	return (int)(2u * (uint32_t)(int)(FBASE / FC4 * f));
}
/* Fast::Increment::operator float  klang.h:4976-4979 */
static float kp_increment_float(int amount) { return kp_bits((uint32_t)((amount >> 9) | 0x3f800000)) - 1.f; }
/* Fast::Phase::operator=(klang::Phase)  klang.h:4993-4997  (2π ↦ 2^31, survey Q2) */
static uint32_t kp_phase_from_radians(float phase) {
	phase = phase * 2147483648.0f / (2.f * KP_PI_F);
	return kp_f2u(phase);
}
/* Fast::Phase::operator float  klang.h:5004-5006 */
static float kp_phase_float(uint32_t position) { return kp_bits((position >> 9) | 0x3f800000) - 1.f; }

/* fast_modp  klang.h:1424-1428 */
static float kp_fast_modp(uint32_t x) { return (kp_bits((x >> 9) | 0x3f800000) - 1.f) * KP_TWO_PI_F; }
/* polysin  klang.h:5093-5096 */
static float kp_polysin(float x) {
	const float x2 = x * x;
	return (((-0.00018542f * x2 + 0.0083143f) * x2 - 0.16666f) * x2 + 1.0f) * x;
}
/* fastsinp  klang.h:5117-5132 */
static float kp_fastsinp(uint32_t p) {
	float x = kp_fast_modp(p);
	if (x > 3.f / 2.f * KP_PI_F) x -= KP_TWO_PI_F;
	else if (x > KP_PI_F / 2.f) x = KP_PI_F - x;
	return kp_polysin(x);
}

/* Fast::Sine  klang.h:5135-5172 */
typedef struct { float frequency; int increment; uint32_t position, offset; } kp_fsine;
static void kp_fsine_init(kp_fsine* s) { s->frequency = 1000.f; s->increment = 0; s->position = 0; s->offset = 0; } /* klang.h:2855 (Q3) */
static void kp_fsine_set_f(kp_fsine* s, float f) { if (f != s->frequency) { s->frequency = f; s->increment = kp_increment_set(f); } }
static void kp_fsine_set_fp(kp_fsine* s, float f, float phase) { s->position = kp_phase_from_radians(phase); s->offset = kp_phase_from_radians(0.f); kp_fsine_set_f(s, f); }
static float kp_fsine_tick(kp_fsine* s) {
	float out = kp_fastsinp(s->position + s->offset);
	s->position += (uint32_t)s->increment;
	return out;
}

/* Fast::OSM  klang.h:5175-5317 */
enum { KP_OSM_SAW = 0, KP_OSM_PULSE = 1 };
typedef struct {
	int waveform;
	int increment; uint32_t offset, duty; int state;
	float delta, f, omf, rcpf, rcpf2, col, c1, c2;
	float frequency;
} kp_osm;

/* OSM::init  klang.h:5206-5215 */
static void kp_osm_init(kp_osm* o) {
	o->state = ((o->offset - (uint32_t)o->increment) < o->duty) ? 3 /*Up*/ : 0 /*Down*/;
	o->f = o->delta;
	o->omf = 1.f - o->f;
	o->rcpf = 1.f / o->f;
	o->rcpf2 = 2.f * o->rcpf;
	o->col = kp_phase_float(o->duty);
	o->c1 = 1.f / o->col;
	o->c2 = -1.f / (1.0f - o->col);
}
/* OSM::setDuty  klang.h:5246-5249 */
static void kp_osm_set_duty(kp_osm* o, float duty) { o->duty = kp_phase_from_radians(duty * (2.f * KP_PI_F)); kp_osm_init(o); }
/* Osm::Osm  klang.h:5323;  Saw duty 0, Triangle 1, Square 1, Pulse .5  klang.h:5348-5354 */
static void kp_osm_construct(kp_osm* o, int waveform, float duty) {
	memset(o, 0, sizeof(*o));
	o->waveform = waveform;
	kp_osm_set_duty(o, duty);
}
static void kp_osm_update_f(kp_osm* o, float frequency) {
	o->frequency = frequency;
	o->increment = kp_increment_set(frequency);
	o->delta = kp_increment_float(o->increment);
}
/* OSM::set(f)  klang.h:5217-5224 */
static void kp_osm_set_f(kp_osm* o, float frequency) { if (o->frequency != frequency) { kp_osm_update_f(o, frequency); kp_osm_init(o); } }
/* OSM::set(f,phase)  klang.h:5226-5234 */
static void kp_osm_set_fp(kp_osm* o, float frequency, float phase) {
	if (o->frequency != frequency) kp_osm_update_f(o, frequency);
	o->offset = kp_phase_from_radians(phase);
	kp_osm_init(o);
}
/* OSM::set(f,phase,duty)  klang.h:5236-5244 */
static void kp_osm_set_fpd(kp_osm* o, float frequency, float phase, float duty) {
	if (o->frequency != frequency) kp_osm_update_f(o, frequency);
	o->offset = kp_phase_from_radians(phase);
	kp_osm_set_duty(o, duty);
}
/* OSM::tick  klang.h:5251-5263 */
static int kp_osm_step(kp_osm* o) {
	o->state = ((o->state << 1) | (o->offset < o->duty ? 1 : 0)) & 3;
	const int transition = o->state | (o->offset < (uint32_t)o->increment ? 4 : 0);
	o->offset += (uint32_t)o->increment;
	return transition;
}
static float kp_sqr(float x) { return x * x; }
/* OSM::saw / OSM::pulse  klang.h:5290-5316.  g++ evaluates tick() before `offset - col` (survey Q1). */
static float kp_osm_tick(kp_osm* o) {
	const int state = kp_osm_step(o);
	const float f = o->f, omf = o->omf, rcpf = o->rcpf, rcpf2 = o->rcpf2, col = o->col, c1 = o->c1, c2 = o->c2;
	if (o->waveform == KP_OSM_SAW) {
		const float p = kp_phase_float(o->offset) - col;
		switch (state) {
		case 3: return c1 * (p + p - f) + 1.f;
		case 0: return c2 * (p + p - f) + 1.f;
		case 2: return rcpf * (c2 * kp_sqr(p) - c1 * kp_sqr(p - f)) + 1.f;
		case 5: return -rcpf * (1.f + c2 * kp_sqr(p + omf) - c1 * kp_sqr(p)) + 1.f;
		case 7: return -rcpf * (1.f + c1 * omf * (p + p + omf)) + 1.f;
		case 4: return -rcpf * (1.f + c2 * omf * (p + p + omf)) + 1.f;
		default: return 0.f;
		}
	} else {
		const float p = kp_phase_float(o->offset);
		switch (state) {
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

/* ====================================== Generic::Oscillator + Generators::Basic */

/* Generic::Oscillator  klang.h:2849-2880;  Phase::operator+=(float)  klang.h:1518-1525 */
typedef struct { float increment, position, frequency, offset, duty; } kp_bosc;
static void kp_bosc_init(kp_bosc* o) { o->increment = 0.f; o->position = 0.f; o->frequency = 1000.f; o->offset = 0.f; o->duty = 0.5f; }
static void kp_bosc_set_f(kp_bosc* o, float f) { o->frequency = f; o->increment = f * 2.f * KP_PI_F / kp_fs.f; }
static void kp_bosc_set_fp(kp_bosc* o, float f, float phase) { o->position = phase; kp_bosc_set_f(o, f); }
static void kp_bosc_advance(kp_bosc* o) {
	if (o->increment >= (2 * KP_PI_F)) return;
	o->position += o->increment;
	if (o->position > (2 * KP_PI_F)) o->position -= (2 * KP_PI_F);
}
enum { KP_B_SINE = 0, KP_B_SAW, KP_B_TRIANGLE, KP_B_SQUARE, KP_B_PULSE };
/* Basic::{Sine,Saw,Triangle,Square,Pulse}::process  klang.h:4899-4944 (sin binds to sinf, survey Q10) */
static float kp_bosc_tick(kp_bosc* o, int wave) {
	float out;
	switch (wave) {
	case KP_B_SINE: out = sinf(o->position + o->offset); break;
	case KP_B_SAW: out = o->position * KP_PI_INV_F - 1.f; break;
	case KP_B_TRIANGLE: out = fabsf(2.f * o->position * KP_PI_INV_F - 2) - 1.f; break;
	case KP_B_SQUARE: out = o->position > KP_PI_F ? 1.f : -1.f; break;
	default: out = o->position > (o->duty * KP_PI_F) ? 1.f : -1.f; break;
	}
	kp_bosc_advance(o);
	return out;
}

/* Wavetable  klang.h:3627-3676; Wavetables::{Sine,Saw}  klang.h:5372-5379 */
#define KP_WT_SIZE 2048
typedef struct { float table[KP_WT_SIZE]; float increment, position, frequency, offset; } kp_wavetable_t;
static void kp_wt_build(kp_wavetable_t* w, int wave) {
	kp_bosc o; kp_bosc_init(&o);
	kp_bosc_set_f(&o, kp_fs.f / KP_WT_SIZE);            /* oscillator.set(fs / size)  klang.h:3646 */
	for (int s = 0; s < KP_WT_SIZE; s++) w->table[s] = kp_bosc_tick(&o, wave);
	w->increment = 0.f; w->position = 0.f; w->frequency = 1000.f; w->offset = 0.f;
}
static void kp_wt_set_f(kp_wavetable_t* w, float f) { w->frequency = f; w->increment = f * (KP_WT_SIZE / kp_fs.f); } /* klang.h:3652-3655 */
static void kp_wt_set_fp(kp_wavetable_t* w, float f, float phase) { w->position = phase * (float)KP_WT_SIZE; kp_wt_set_f(w, f); }
/* buffer::operator[](float)  klang.h:2070-2078 */
static float kp_lerp_table(const float* samples, int size, float offset) {
	const float f = floorf(offset);
	const float frac = offset - f;
	const int i = (int)offset;
	const int j = (i == (size - 1)) ? 0 : (i + 1);
	return samples[i] * (1.f - frac) + samples[j] * frac;
}
/* Wavetable::process  klang.h:3672-3675; Phase::operator+=(const increment&)  klang.h:1527-1534 */
static float kp_wt_tick(kp_wavetable_t* w) {
	const float size = (float)KP_WT_SIZE;
	if (!(w->increment >= size)) {
		w->position += w->increment;
		if (w->position > size) w->position -= size;
	}
	return kp_lerp_table(w->table, KP_WT_SIZE, w->position + w->offset);
}

/* ============================================================ Filters (5383-5813) */

enum { KP_BQ_LPF = 0, KP_BQ_HPF, KP_BQ_BPF, KP_BQ_BRF, KP_BQ_APF, KP_BQ_BW2 };
/* Biquad::Filter  klang.h:5550-5612 */
typedef struct { int type; float f, Q, a1, a2, b0, b1, b2, a, cos0, sin0, z0, z1, in, out; } kp_biquad;
static void kp_biquad_construct(kp_biquad* b, int type) {
	memset(b, 0, sizeof(*b));
	b->type = type; b->b0 = 1.f; b->cos0 = 1.f;
}
/* Filter::reset  klang.h:5565-5572 */
static void kp_biquad_reset(kp_biquad* b) {
	b->f = 0; b->Q = 0; b->b0 = 1; b->a1 = b->a2 = b->b1 = b->b2 = 0; b->a = 0; b->z0 = b->z1 = 0;
}
/* constant{x}.inv  klang.h:96-98 : float(1.0 / double(x)) */
static float kp_const_inv(float x) { const double v = (double)x; return v == 0.0 ? 0.0f : (float)(1.0 / v); }
/* LPF::init 5658-5665, HPF::init 5675-5682, BPF::init_peak 5720-5729, BRF::init 5734-5739, Butterworth::LPF<2> 5803-5810 */
static void kp_biquad_init(kp_biquad* b) {
	const float inv = kp_const_inv(1.f + b->a);
	const float cos0 = b->cos0, a = b->a;
	switch (b->type) {
	case KP_BQ_LPF:
		b->a1 = inv * (-2.f * cos0); b->a2 = inv * (1.f - a);
		b->b2 = b->b0 = inv * (1.f - cos0) * 0.5f; b->b1 = inv * (1.f - cos0); break;
	case KP_BQ_HPF:
		b->a1 = inv * (-2.f * cos0); b->a2 = inv * (1.f - a);
		b->b2 = b->b0 = inv * (1.f + cos0) * 0.5f; b->b1 = inv * -(1.f + cos0); break;
	case KP_BQ_BPF:
		b->a1 = inv * (-2.f * cos0); b->a2 = inv * (1.f - a);
		b->b0 = inv * a; b->b1 = 0; b->b2 = inv * -a; break;
	case KP_BQ_BRF:
		b->b1 = b->a1 = inv * (-2.f * cos0); b->a2 = inv * (1.f - a); b->b0 = b->b2 = inv; break;
	case KP_BQ_BW2:
		b->b0 = inv * ((1.f - cos0) / 2.f); b->b1 = inv * (1.f - cos0); b->b2 = inv * ((1.f - cos0) / 2.f);
		b->a1 = inv * (-2.f * cos0); b->a2 = inv * (1.f - a); break;
	default: break;
	}
}
/* Filter::set(f,Q)  klang.h:5584-5600 */
static void kp_biquad_set(kp_biquad* b, float f, float Q) {
	if (Q < 0) Q = f / -Q;
	if (b->f != f || b->Q != Q) {
		b->f = f; b->Q = Q;
		const float w = f * kp_fs.w;
		b->cos0 = cosf(w);
		b->sin0 = sinf(w);
		if (Q < 0.5) Q = 0.5f;
		b->a = b->sin0 / (2.f * Q);
		kp_biquad_init(b);
	}
}
/* APF::set(f,r) + APF::init  klang.h:5752-5772 (cos evaluated in double on a float omega) */
static void kp_apf_set(kp_biquad* b, float f, float r) {
	if (b->f != f || b->a != r) {
		b->f = f; b->a = r;
		const float w = f * kp_fs.w;
		b->cos0 = cosf(w); b->sin0 = sinf(w);
		float omega = 2.0f * KP_PI_F * f / kp_fs.f;
		float c0 = cosf(omega);
		b->b0 = b->a2 = b->a * b->a;
		b->b1 = b->a1 = (-2.f * b->a * c0);
		b->b2 = 1.f;
	}
}
/* Filter::set(f)  klang.h:5575 */
static void kp_biquad_set_f(kp_biquad* b, float f) {
	if (b->type == KP_BQ_APF) kp_apf_set(b, f, 1.f); else kp_biquad_set(b, f, KP_ROOT2_INV_F);
}
/* Filter::process  klang.h:5605-5612 */
static float kp_biquad_tick(kp_biquad* b, float in) {
	const float z0 = b->z0, z1 = b->z1;
	const float y = b->b0 * in + z0;
	b->z0 = b->b1 * in - b->a1 * y + z1;
	b->z1 = b->b2 * in - b->a2 * y;
	b->in = in; b->out = y;
	return y;
}

enum { KP_OP_LPF = 0, KP_OP_HPF, KP_OP_BW1 };
/* OnePole::Filter  klang.h:5470-5543, Butterworth::LPF<1>  klang.h:5786-5799 */
typedef struct { int type; float f, a1, b0, b1, z, out; } kp_onepole;
static void kp_onepole_construct(kp_onepole* p, int type) { memset(p, 0, sizeof(*p)); p->type = type; p->b0 = 1.f; }
static void kp_onepole_reset(kp_onepole* p) { p->a1 = 0; p->b0 = 1; p->b1 = 0; p->f = 0; p->z = 0; } /* klang.h:5482-5488 */
static void kp_onepole_set(kp_onepole* p, float f) {
	if (p->f != f) {
		p->f = f;
		if (p->type == KP_OP_LPF) { const float e = expf(-f * kp_fs.w); p->b0 = 1 - e; p->a1 = e; }
		else if (p->type == KP_OP_HPF) { const float e = expf(-f * kp_fs.w); p->b0 = 0.5f * (1.f + e); p->b1 = -p->b0; p->a1 = e; }
		else { const float c = 1.f / tanf(KP_PI_F * f * kp_fs.inv); const float inv = kp_const_inv(1.f + c); p->b0 = inv; p->a1 = (1.f - c) * inv; }
	}
}
static float kp_onepole_tick(kp_onepole* p, float in) {
	if (p->type == KP_OP_LPF) p->out = p->b0 * in + p->a1 * p->out + KP_DENORMALISE;               /* klang.h:5515-5517 */
	else if (p->type == KP_OP_HPF) { p->out = p->b0 * in + p->b1 * p->z + p->a1 * p->out + KP_DENORMALISE; p->z = in; } /* klang.h:5499-5502 */
	else { p->out = p->b0 * (in + p->z) - p->a1 * p->out; p->z = in; }                           /* klang.h:5795-5798 */
	return p->out;
}

/* ========================================================== Envelope (3723-4137) */

#define KP_ENV_MAXPTS 16
enum { KP_ENV_SUSTAIN = 0, KP_ENV_RELEASE = 1, KP_ENV_OFF = 2 };
typedef struct {
	float px[KP_ENV_MAXPTS], py[KP_ENV_MAXPTS]; int npoints;
	int loop_start, loop_end;
	int point; float time, timeInc; int stage;
	float out;                         /* Envelope::out */
	float r_out, r_target, r_rate; int r_active; /* Envelope::Linear ramp */
	float A, D, S, R;                  /* ADSR params */
} kp_env;

static void kp_ramp_set_value(kp_env* e, float v) { e->r_out = v; e->r_target = v; e->r_active = 0; }   /* klang.h:3763-3767 */
static void kp_ramp_set_target(kp_env* e, float t) { e->r_target = t; e->r_active = (e->r_out != t); }   /* klang.h:3757-3760 */
/* Envelope::setTargetTime  klang.h:4077-4081 (abs ≡ fabsf, survey Q4) */
static void kp_env_set_target(kp_env* e, float x, float y, float time) {
	e->time = time;
	kp_ramp_set_target(e, y);
	e->r_rate = fabsf(y - e->r_out) / ((x - time) * kp_fs.f);
}
/* Envelope::initialise  klang.h:3974-3989 */
static void kp_env_initialise(kp_env* e) {
	e->point = 0;
	e->timeInc = 1.0f / kp_fs.f;
	e->loop_start = e->loop_end = -1;
	e->stage = KP_ENV_SUSTAIN;
	if (e->npoints) {
		e->out = e->py[0];
		kp_ramp_set_value(e, e->py[0]);
		if (e->npoints > 1) kp_env_set_target(e, e->px[1], e->py[1], e->px[0]);
	} else {
		e->out = 1.0f;
		kp_ramp_set_value(e, 1.0f);
	}
}
/* Envelope::Envelope()  klang.h:3867: one point (0,1) */
static void kp_env_construct(kp_env* e) {
	memset(e, 0, sizeof(*e));
	e->r_rate = 0.f;
	e->npoints = 1; e->px[0] = 0.f; e->py[0] = 1.f;
	kp_env_initialise(e);
}
/* Envelope::set(points) / operator=(initializer_list)  klang.h:3887-3896 */
static void kp_env_set_points(kp_env* e, int n, const float* xy) {
	e->npoints = n;
	for (int p = 0; p < n; p++) { e->px[p] = xy[2 * p]; e->py[p] = xy[2 * p + 1]; }
	kp_env_initialise(e);
}
/* Envelope::setLoop  klang.h:3923-3926 */
static void kp_env_set_loop(kp_env* e, int s, int t) { if (s >= 0 && t < e->npoints) { e->loop_start = s; e->loop_end = t; } }
/* Envelope::release  klang.h:3961-3966 */
static void kp_env_release(kp_env* e, float time, float level) { e->stage = KP_ENV_RELEASE; kp_env_set_target(e, time, level, 0.f); }
/* Envelope::process  klang.h:4018-4051 with Linear::operator++  klang.h:3785-3806 */
static float kp_env_tick(kp_env* e) {
	const float output = e->r_out;
	if (e->r_active) {
		if (e->r_target > e->r_out) {
			e->r_out += e->r_rate;
			if (e->r_out >= e->r_target) { e->r_out = e->r_target; e->r_active = 0; }
		} else {
			e->r_out -= e->r_rate;
			if (e->r_out <= e->r_target) { e->r_out = e->r_target; e->r_active = 0; }
		}
	}
	e->out = output;
	switch (e->stage) {
	case KP_ENV_SUSTAIN:
		e->time += e->timeInc;
		if (!e->r_active) {
			const int loop_active = e->loop_start != -1 && e->loop_end != -1;
			if (loop_active && (e->point + 1) >= e->loop_end) {
				e->point = e->loop_start;
				kp_ramp_set_value(e, e->py[e->point]);
				if (e->loop_start != e->loop_end)
					kp_env_set_target(e, e->px[e->point + 1], e->py[e->point + 1], e->px[e->point]);
			} else if ((e->point + 1) < e->npoints) {
				if (e->time >= e->px[e->point + 1]) {
					e->point++;
					kp_ramp_set_value(e, e->py[e->point]);
					if ((e->point + 1) < e->npoints)
						kp_env_set_target(e, e->px[e->point + 1], e->py[e->point + 1], e->px[e->point]);
				}
			} else {
				e->stage = KP_ENV_OFF;
			}
		}
		break;
	case KP_ENV_RELEASE:
		if (!e->r_active) e->stage = KP_ENV_OFF;
		break;
	default: break;
	}
	return e->out;
}
/* Envelope::at  klang.h:3929-3942 */
static float kp_env_at(const kp_env* e, float time) {
	if (e->npoints == 0) return 0;
	float lx = 0, ly = e->py[0];
	for (int p = 0; p < e->npoints; p++) {
		if (e->px[p] >= time) {
			const float dx = e->px[p] - lx;
			const float dy = e->py[p] - ly;
			const float x = time - lx;
			return dx == 0 ? ly : (ly + x * dy / dx);
		}
		lx = e->px[p]; ly = e->py[p];
	}
	return e->py[e->npoints - 1];
}
/* ADSR::set  klang.h:4115-4128;  ADSR::ADSR  klang.h:4113 */
static void kp_adsr_set(kp_env* e, float attack, float decay, float sustain, float release) {
	e->A = attack; e->D = decay + 0.005f; e->S = sustain; e->R = release + 0.005f;
	e->npoints = 3;
	e->px[0] = 0; e->py[0] = 0;
	e->px[1] = e->A; e->py[1] = 1;
	e->px[2] = e->A + e->D; e->py[2] = e->S;
	kp_env_initialise(e);
	kp_env_set_loop(e, 2, 2);
}
static void kp_adsr_construct(kp_env* e) { kp_env_construct(e); kp_adsr_set(e, 0.5f, 0.5f, 1.f, 0.5f); }
/* ADSR::release  klang.h:4130-4132 */
static void kp_adsr_release(kp_env* e) { kp_env_release(e, e->R, 0.f); }

/* ============================================================ Delay (3381-3512) */

typedef struct { float* buf; int SIZE; float time; int position; int last_position; float last_fraction; float out; } kp_delay;
static void kp_delay_construct(kp_delay* d, int size) {
	d->buf = (float*)calloc((size_t)size + 1, sizeof(float));
	d->SIZE = size; d->time = 1; d->position = 0; d->last_position = 0; d->last_fraction = 0.f; d->out = 0.f;
}
static void kp_delay_free(kp_delay* d) { free(d->buf); d->buf = NULL; }
/* Delay::input  klang.h:3396-3403 */
static void kp_delay_write(kp_delay* d, float in) {
	d->buf[d->position] = in;
	d->position++;
	if (d->position == d->SIZE) d->position = 0;
}
/* Delay::tap(int)  klang.h:3405-3410 */
static float kp_delay_tap_i(const kp_delay* d, int delay) {
	int read = (d->position - 1) - delay;
	if (read < 0) read += d->SIZE;
	return d->buf[read];
}
/* Delay::tap(float)  klang.h:3412-3427 */
static float kp_delay_tap_f(const kp_delay* d, float delay) {
	float read = (float)(d->position - 1) - delay;
	if (read < 0.f) read += d->SIZE;
	const int i = (int)read;
	const float fraction = read - i;
	const int j = (i + 1) % d->SIZE;
	return d->buf[i] + fraction * (d->buf[j] - d->buf[i]);
}
/* Delay::lagrange  klang.h:3429-3458 */
static float kp_delay_lagrange(const kp_delay* d, float delay) {
	float read = (float)(d->position - 1) - delay;
	if (read < 0.f) read += d->SIZE;
	int i = (int)read;
	float x = read - i;
	int i0 = (i - 1 + d->SIZE) % d->SIZE, i1 = i, i2 = (i + 1) % d->SIZE, i3 = (i + 2) % d->SIZE;
	float y0 = d->buf[i0], y1 = d->buf[i1], y2 = d->buf[i2], y3 = d->buf[i3];
	float c0 = (-x * (x - 1) * (x - 2)) / 6.0f;
	float c1 = ((x + 1) * (x - 1) * (x - 2)) / 2.0f;
	float c2 = (-x * (x + 1) * (x - 2)) / 2.0f;
	float c3 = (x * (x + 1) * (x - 1)) / 6.0f;
	return c0 * y0 + c1 * y1 + c2 * y2 + c3 * y3;
}
/* Delay::set  klang.h:3480-3489 */
static void kp_delay_set(kp_delay* d, float samples) {
	d->time = samples < d->SIZE ? (float)samples : d->SIZE;
	float read = (float)(d->position - 1) - d->time;
	if (read < 0.f) read += d->SIZE;
	d->last_position = (int)read;
	d->last_fraction = read - d->last_position;
}
/* Delay::process  klang.h:3461-3473 */
static float kp_delay_tick(kp_delay* d) {
	const int i = d->last_position;
	const int j = (i + 1) % d->SIZE;
	d->out = d->buf[i] + d->last_fraction * (d->buf[j] - d->buf[i]);
	d->last_position = (d->last_position + 1) % d->SIZE;
	return d->out;
}
/* Stereo::Delay::tap(float)  klang.h:4668-4681: both channels read at items[0].position */
static void kp_sdelay_tap_f(const kp_delay* l, const kp_delay* r, float delay, float* ol, float* orr) {
	float read = (float)(l->position - 1) - delay;
	if (read < 0.f) read += l->SIZE;
	const float f = floorf(read);
	delay = read - f;
	const int i = (int)read;
	const int j = (i == (l->SIZE - 1)) ? 0 : (i + 1);
	*ol = l->buf[i] * (1.f - delay) + l->buf[j] * delay;
	*orr = r->buf[i] * (1.f - delay) + r->buf[j] * delay;
}

/* ======================================================== primitive KAT runners */

enum { OSC_FAST_SAW = 0, OSC_FAST_TRIANGLE, OSC_FAST_SQUARE, OSC_FAST_PULSE, OSC_FAST_SINE,
       OSC_BASIC_SINE, OSC_BASIC_SAW, OSC_BASIC_TRIANGLE, OSC_BASIC_SQUARE, OSC_BASIC_PULSE,
       OSC_WT_SINE, OSC_WT_SAW, OSC_BASIC_NOISE, OSC_FAST_NOISE };

int kp_osc(int kind, int nargs, float f, float phase, float duty, int n, float* out) {
	kp_ensure_fs();
	if (kind <= OSC_FAST_PULSE) {
		static const int wf[4] = { KP_OSM_SAW, KP_OSM_SAW, KP_OSM_PULSE, KP_OSM_PULSE };
		static const float dt[4] = { 0.f, 1.f, 1.f, 0.5f };
		kp_osm o; kp_osm_construct(&o, wf[kind], dt[kind]);
		if (nargs == 1) kp_osm_set_f(&o, f); else if (nargs == 2) kp_osm_set_fp(&o, f, phase); else kp_osm_set_fpd(&o, f, phase, duty);
		for (int s = 0; s < n; s++) out[s] = kp_osm_tick(&o);
	} else if (kind == OSC_FAST_SINE) {
		kp_fsine o; kp_fsine_init(&o);
		if (nargs == 1) kp_fsine_set_f(&o, f); else if (nargs == 2) kp_fsine_set_fp(&o, f, phase);
		for (int s = 0; s < n; s++) out[s] = kp_fsine_tick(&o);
	} else if (kind <= OSC_BASIC_PULSE) {
		kp_bosc o; kp_bosc_init(&o);
		if (nargs == 1) kp_bosc_set_f(&o, f); else if (nargs == 2) kp_bosc_set_fp(&o, f, phase);
		else if (kind == OSC_BASIC_PULSE) { kp_bosc_set_fp(&o, f, phase); o.duty = duty; }   /* klang.h:4935-4938 */
		for (int s = 0; s < n; s++) out[s] = kp_bosc_tick(&o, kind - OSC_BASIC_SINE);
	} else if (kind <= OSC_WT_SAW) {
		kp_wavetable_t* w = (kp_wavetable_t*)malloc(sizeof(kp_wavetable_t));
		kp_wt_build(w, kind == OSC_WT_SINE ? KP_B_SINE : KP_B_SAW);
		if (nargs == 1) kp_wt_set_f(w, f); else if (nargs == 2) kp_wt_set_fp(w, f, phase);
		for (int s = 0; s < n; s++) out[s] = kp_wt_tick(w);
		free(w);
	} else if (kind == OSC_BASIC_NOISE) {          /* Generators::Basic::Noise::process  klang.h:4947-4951: one libc rand() per tick */
		for (int s = 0; s < n; s++) out[s] = rand() * 2.f / (const float)RAND_MAX - 1.f;
	} else if (kind == OSC_FAST_NOISE) {           /* Generators::Fast::Noise::process   klang.h:5357-5366 */
		for (int s = 0; s < n; s++) {
			union { unsigned int i; float f; } u;
			u.i = ((rand() & 0x7fffu) << 1) | 0x43800000u;
			out[s] = u.f - 257.f;
		}
	} else return -1;
	return 0;
}

int kp_wavetable(int kind, float* table) {
	kp_ensure_fs();
	if (kind != OSC_WT_SINE && kind != OSC_WT_SAW) return -1;
	kp_wavetable_t* w = (kp_wavetable_t*)malloc(sizeof(kp_wavetable_t));
	kp_wt_build(w, kind == OSC_WT_SINE ? KP_B_SINE : KP_B_SAW);
	memcpy(table, w->table, sizeof(w->table));
	free(w);
	return 0;
}

enum { FLT_BIQUAD_LPF = 0, FLT_BIQUAD_HPF, FLT_ONEPOLE_LPF, FLT_ONEPOLE_HPF,
       FLT_BIQUAD_BPF, FLT_BIQUAD_BRF, FLT_BIQUAD_APF, FLT_BUTTERWORTH_LPF1, FLT_BUTTERWORTH_LPF2,
       FLT_DCF, FLT_IIR1, FLT_IIR2, FLT_MODAL, FLT_FOLLOWER_PEAK, FLT_FOLLOWER_RMS, FLT_WINDOW_MEAN, FLT_WINDOW_RMS };

int kp_filter(int kind, int nset, const float* f, const float* Q, int n, const float* in, float* out, float* coeffs) {
	kp_ensure_fs();
	int bq = -1, op = -1;
	if (kind == FLT_DCF) {                       /* Filters::DCF  klang.h:5387-5398: out = in - z + r*out; z = in */
		float r = 0.995f, z = 0.f, o = 0.f;
		for (int s = 0; s < n; s++) { if (s < nset) r = f[s]; o = in[s] - z + r * o; z = in[s]; out[s] = o; }
		if (coeffs) { coeffs[0] = r; coeffs[1] = z; coeffs[2] = coeffs[3] = coeffs[4] = 0; }
		return 0;
	}
	if (kind == FLT_IIR1) {                      /* Filters::IIR<1>  klang.h:5433-5446: out = in*a + out*b */
		float a = 1.f, b = 0.f, o = 0.f;
		for (int s = 0; s < n; s++) { if (s < nset) { a = f[s]; b = 1.f - a; } o = in[s] * a + o * b; out[s] = o; }
		if (coeffs) { coeffs[0] = a; coeffs[1] = b; coeffs[2] = coeffs[3] = coeffs[4] = 0; }
		return 0;
	}
	if (kind == FLT_IIR2) {                      /* Filters::IIR<2>  klang.h:5401-5430 */
		if (!Q) return -1;
		float a[2] = { 0.f, 0.f }, y[2] = { 0.f, 0.f };
		for (int s = 0; s < n; s++) {
			if (s < nset) { a[0] = f[s]; a[1] = Q[s]; }
			float o = in[s];
			o -= a[0] * y[0];
			o -= a[1] * y[1];
			y[1] = y[0]; y[0] = o;
			out[s] = o;
		}
		if (coeffs) { coeffs[0] = a[0]; coeffs[1] = a[1]; coeffs[2] = y[0]; coeffs[3] = y[1]; coeffs[4] = 0; }
		return 0;
	}
	if (kind == FLT_MODAL) {                     /* Modifiers::Modal  klang.h:5817-5857 */
		if (!Q) return -1;
		float a1 = 0.f, a2 = 0.f, y1 = 0.f, y2 = 0.f, gain = 0.05f;
		for (int s = 0; s < n; s++) {
			if (s < nset) {                      /* set(f, decay): exp / cos bind to expf / cosf (float arguments) */
				gain = 0.05f;
				const float w = f[s] * kp_fs.w;
				float d = expf(-KP_PI_F / (Q[s] * kp_fs.f));
				d = d < 1e-6f ? 1e-6f : (0.9999f < d ? 0.9999f : d);
				float c = cosf(w);
				c = c < -0.9999f ? -0.9999f : (0.9999f < c ? 0.9999f : c);
				a1 = 2.f * d * c;
				a2 = -d * d;
				y2 = 0.f; y1 = 0.f;
			}
			const float x = in[s] * gain;        /* input(): in *= gain */
			const float o = x + a1 * y1 + a2 * y2;
			y2 = y1; y1 = o;
			out[s] = o;
		}
		if (coeffs) { coeffs[0] = a1; coeffs[1] = a2; coeffs[2] = gain; coeffs[3] = y1; coeffs[4] = y2; }
		return 0;
	}
	if (kind == FLT_FOLLOWER_PEAK || kind == FLT_FOLLOWER_RMS) {   /* Envelope::Follower + AR  klang.h:5862-5896 */
		if (!Q) return -1;
		float attack = 0.01f, release = 0.1f;    /* Follower() { set(0.01f, 0.1f); } */
		float A = 1.f - expf(-1.0f / (kp_fs.f * attack)), R = 1.f - expf(-1.0f / (kp_fs.f * release)), ar = 0.f;
		for (int s = 0; s < n; s++) {
			if (s < nset && (attack != f[s] || release != Q[s])) {
				attack = f[s]; release = Q[s];
				A = 1.f - (attack == 0.f ? 0.f : expf(-1.0f / (kp_fs.f * attack)));
				R = 1.f - (release == 0.f ? 0.f : expf(-1.0f / (kp_fs.f * release)));
			}
			const float x = kind == FLT_FOLLOWER_PEAK ? fabsf(in[s]) : in[s] * in[s];
			const float smoothing = x > ar ? A : R;
			ar = ar + smoothing * (x - ar);
			out[s] = kind == FLT_FOLLOWER_PEAK ? ar : sqrtf(ar);
		}
		if (coeffs) { coeffs[0] = A; coeffs[1] = R; coeffs[2] = ar; coeffs[3] = coeffs[4] = 0; }
		return 0;
	}
	if (kind == FLT_WINDOW_MEAN || kind == FLT_WINDOW_RMS) {   /* Envelope::Follower::Window<64>  klang.h:5904-5948 */
		if (!Q) return -1;
		enum { WINDOW = 64 };
		const float inv = (float)(1.0 / (double)WINDOW);         /* constant window = { WINDOW }; window.inv  klang.h:93-104 */
		float buf[WINDOW] = { 0 };
		int pos = 0;
		double sum = 0;                                           /* 64-bit on purpose (klang.h:5909) */
		float attack = 0.01f, release = 0.1f;                     /* Window() { set(0.01f, 0.1f); } */
		float A = 1.f - expf(-1.0f / (kp_fs.f * attack)), R = 1.f - expf(-1.0f / (kp_fs.f * release)), ar = 0.f;
		for (int s = 0; s < n; s++) {
			if (s < nset && (attack != f[s] || release != Q[s])) {
				attack = f[s]; release = Q[s];
				A = 1.f - (attack == 0.f ? 0.f : expf(-1.0f / (kp_fs.f * attack)));
				R = 1.f - (release == 0.f ? 0.f : expf(-1.0f / (kp_fs.f * release)));
			}
			sum -= (double)buf[pos];
			buf[pos] = kind == FLT_WINDOW_MEAN ? fabsf(in[s]) : in[s] * in[s];
			sum += (double)buf[pos];
			if (++pos == WINDOW) pos = 0;
			float x = (float)(sum * inv);
			if (kind == FLT_WINDOW_RMS) x = sqrtf(x);
			const float smoothing = x > ar ? A : R;
			ar = ar + smoothing * (x - ar);
			out[s] = ar;
		}
		if (coeffs) { coeffs[0] = A; coeffs[1] = R; coeffs[2] = ar; coeffs[3] = (float)sum; coeffs[4] = 0; }
		return 0;
	}
	switch (kind) {
	case FLT_BIQUAD_LPF: bq = KP_BQ_LPF; break;
	case FLT_BIQUAD_HPF: bq = KP_BQ_HPF; break;
	case FLT_BIQUAD_BPF: bq = KP_BQ_BPF; break;
	case FLT_BIQUAD_BRF: bq = KP_BQ_BRF; break;
	case FLT_BIQUAD_APF: bq = KP_BQ_APF; break;
	case FLT_BUTTERWORTH_LPF2: bq = KP_BQ_BW2; break;
	case FLT_ONEPOLE_LPF: op = KP_OP_LPF; break;
	case FLT_ONEPOLE_HPF: op = KP_OP_HPF; break;
	case FLT_BUTTERWORTH_LPF1: op = KP_OP_BW1; break;
	default: return -1;
	}
	if (bq >= 0) {
		kp_biquad b; kp_biquad_construct(&b, bq);
		for (int s = 0; s < n; s++) {
			if (s < nset) {
				if (Q) { if (bq == KP_BQ_APF) kp_apf_set(&b, f[s], Q[s]); else kp_biquad_set(&b, f[s], Q[s]); }
				else kp_biquad_set_f(&b, f[s]);
			}
			out[s] = kp_biquad_tick(&b, in[s]);
		}
		if (coeffs) { coeffs[0] = b.b0; coeffs[1] = b.b1; coeffs[2] = b.b2; coeffs[3] = b.a1; coeffs[4] = b.a2; }
	} else {
		kp_onepole p; kp_onepole_construct(&p, op);
		for (int s = 0; s < n; s++) {
			if (s < nset) kp_onepole_set(&p, f[s]);
			out[s] = kp_onepole_tick(&p, in[s]);
		}
		if (coeffs) { coeffs[0] = p.b0; coeffs[1] = p.b1; coeffs[2] = 0; coeffs[3] = p.a1; coeffs[4] = 0; }
	}
	return 0;
}

int kp_envelope(int npts, const float* xy, int loop_start, int loop_end, int n, int release_at,
                float release_time, float release_level, float* out, int* stage_out) {
	kp_ensure_fs();
	if (npts > KP_ENV_MAXPTS) return -1;
	kp_env e; kp_env_construct(&e);
	kp_env_set_points(&e, npts, xy);
	if (loop_start >= 0) kp_env_set_loop(&e, loop_start, loop_end);
	for (int s = 0; s < n; s++) {
		if (s == release_at) kp_env_release(&e, release_time, release_level);
		out[s] = kp_env_tick(&e);
		if (stage_out) stage_out[s] = e.stage;
	}
	return 0;
}

int kp_envelope_at(int npts, const float* xy, int n, const float* t, float* out) {
	kp_ensure_fs();
	if (npts > KP_ENV_MAXPTS) return -1;
	kp_env e; kp_env_construct(&e);
	kp_env_set_points(&e, npts, xy);
	for (int s = 0; s < n; s++) out[s] = kp_env_at(&e, t[s]);
	return 0;
}

int kp_adsr(float A, float D, float S, float R, int n, int release_at, float* out, int* stage_out) {
	kp_ensure_fs();
	kp_env e; kp_adsr_construct(&e);
	kp_adsr_set(&e, A, D, S, R);
	for (int s = 0; s < n; s++) {
		if (s == release_at) kp_adsr_release(&e);
		out[s] = kp_env_tick(&e);
		if (stage_out) stage_out[s] = e.stage;
	}
	return 0;
}

int kp_delay1000(int n, const float* in, const int* di, const float* df, const float* set_at,
                 float* out_i, float* out_f, float* out_p) {
	kp_delay d; kp_delay_construct(&d, 1000);
	int have_set = 0;
	for (int s = 0; s < n; s++) {
		kp_delay_write(&d, in[s]);
		out_i[s] = kp_delay_tap_i(&d, di[s]);
		out_f[s] = kp_delay_tap_f(&d, df[s]);
		if (set_at[s] >= 0.f) { kp_delay_set(&d, set_at[s]); have_set = 1; }
		out_p[s] = have_set ? kp_delay_tick(&d) : 0.f;
	}
	kp_delay_free(&d);
	return 0;
}

int kp_delay1000_lagrange(int n, const float* in, const float* df, float* out) {
	kp_delay d; kp_delay_construct(&d, 1000);
	for (int s = 0; s < n; s++) { kp_delay_write(&d, in[s]); out[s] = kp_delay_lagrange(&d, df[s]); }
	kp_delay_free(&d);
	return 0;
}

int kp_stereo_delay1000(int n, const float* inl, const float* inr, const float* df, float* outl, float* outr) {
	kp_delay l, r; kp_delay_construct(&l, 1000); kp_delay_construct(&r, 1000);
	for (int s = 0; s < n; s++) {
		kp_delay_write(&l, inl[s]); kp_delay_write(&r, inr[s]);
		kp_sdelay_tap_f(&l, &r, df[s], &outl[s], &outr[s]);
	}
	kp_delay_free(&l); kp_delay_free(&r);
	return 0;
}

int kp_control_smooth(float lo, float hi, float initial, int n, const float* values, float* out) {
	kp_control c = kp_dial(lo, hi, initial);
	for (int s = 0; s < n; s++) { kp_control_set(&c, values[s]); out[s] = kp_control_smooth1(&c); }
	return 0;
}

/* the graph layer lives in klang_port_graphs.inc to keep this file readable */
#include "klang_port_graphs.inc"
