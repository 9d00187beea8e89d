// klang-b200 — the user graphs of the BASELINE configs over the primitives of kb_prims.cuh.
//
// For each graph: the HOST half = constructor, Note::on()/off(), Effect::prepare() (event / block rate; draws
// libc rand() and uses the host libm like the reference), and the DEVICE half = process() (per sample).
#pragma once
#include <stdlib.h>

#include "kb_prims.cuh"

#include "../../include/klang_b200.h"   // graph ids KB_SY_* / KB_FX_*
#define KB_SY_COUNT 15
#define KB_FX_COUNT 18

// =========================================================================================== HOST halves

// random<float>(min,max) / random<double>                                    klang.h:236-237
inline float kb_randomf(float lo, float hi) { return rand() * ((hi - lo) / (float)RAND_MAX) + lo; }
inline double kb_randomd(double lo, double hi) { return rand() * ((hi - lo) / (double)RAND_MAX) + lo; }

// power(float base, float exp)                                               klang.h:187-218
inline float kb_powerf(float base, float e) {
	if (base == 10.f) return (float)::expf(e * (float)2.3025850929940456840179914546843642076011014886287729760333279009);
	else if (e == 0.f) return 1.f;
	else if (e == 1.f) return base;
	else if (e == 2.f) return base * base;
	else if (e == 3.f) return base * base * base;
	else if (e == 4.f) return base * base * base * base;
	else if (e == -1.f) return 1.f / base;
	else if (e == -2.f) return 1.f / (base * base);
	else if (e == -3.f) return 1.f / (base * base * base);
	else if (e == -4.f) return 1.f / (base * base * base * base);
	return ::powf(base, e);
}
// Pitch::operator-> Frequency                                                klang.h:1568-1571
inline float kb_pitch_to_frequency_host(float p) { return 440.f * kb_powerf(2.f, (p - 69.f) / 12.f); }
inline KbControl kb_dial(float lo, float hi, float initial) { KbControl c = { lo, hi, initial, 0.f }; return c; }   // klang.h:1796-1799

// ---- Subtractive / Filter.k (examples/Subtractive/Filter.k:7-36; canonical C2 graph = Saw osc, ADSR from controls 0..3)
inline void kb_sub_construct(const KbFs& fs, int graph, KbSubVoice& n) {
	if (graph == KB_SY_FILTER_K) kb_osm_construct(n.osc, 1, 1.0f); else kb_osm_construct(n.osc, 0, 0.f);
	kb_adsr_construct(fs, n.adsr);
	kb_env_construct(fs, n.env);
	kb_biquad_construct(n.filter, KB_BQ_LPF);
}
inline void kb_sub_on(const KbFs& fs, int graph, const KbControl* c, KbSubVoice& n, float pitch) {
	const float f = kb_pitch_to_frequency_host(pitch);                               // Filter.k:16
	kb_osm_set_fp(fs, n.osc, f, 0.f);                                           // Filter.k:17
	if (graph == KB_SY_SUBTRACTIVE) kb_adsr_set(fs, n.adsr, c[0].value, c[1].value, c[2].value, c[3].value);
	else kb_adsr_set(fs, n.adsr, 0.f, 0.f, 1.f, 0.25f);                         // Filter.k:19
	const float pts[6] = { 0.f, f * 2, 0.25f, f * 10, 2.f, f * 5 };             // Filter.k:20
	kb_env_set_points(fs, n.env, 3, pts);
	kb_biquad_reset(n.filter);                                                  // Filter.k:22
}

// ---- SuperSaw (examples/SuperSaw.k:7-34)
inline void kb_ssaw_construct(const KbFs& fs, KbSsawVoice& n) {
	for (int k = 0; k < 7; k++) kb_osm_construct(n.osc[k], 0, 0.f);
	kb_adsr_construct(fs, n.adsr);
}
inline void kb_ssaw_on(const KbFs& fs, const KbControl* c, KbSsawVoice& n, float pitch) {
	const float f = kb_pitch_to_frequency_host(pitch);                                           // SuperSaw.k:13
	const float detune = (float)(0.01 * (double)c[2].value * (double)f);                    // SuperSaw.k:14
	for (int k = 0; k < 7; k++) {
		const double rnd = kb_randomd(0.999, 1.001);
		const float fk = f + (float)((double)((float)(k - 3) * detune) * rnd);              // SuperSaw.k:17
		kb_osm_set_fpd(fs, n.osc[k], fk, 0.f, c[1].value);
	}
	kb_adsr_set(fs, n.adsr, c[0].value, 0.25f, 1.0f, 0.5f);                                 // SuperSaw.k:18
}

// ---- FM (examples/FM.k:27-74)
inline void kb_fm_construct(const KbFs& fs, KbFmVoice& n) {
	for (int k = 0; k < 3; k++) { kb_fsine_init(n.op[k].osc); kb_env_construct(fs, n.op[k].env); n.op[k].amp = 1.f; n.op[k].in = 0.f; }
	kb_adsr_construct(fs, n.adsr);
}
inline void kb_fm_on(const KbFs& fs, const KbControl* c, KbFmVoice& n, float pitch) {
	const float fc = kb_pitch_to_frequency_host(pitch);                                     // FM.k:42
	const float fd = fc * c[0].value;                                                       // FM.k:43
	const float e1[4] = { 0.f, 0.f, 3.f, 1.f }, e2[4] = { 0.f, 1.5f, 3.f, 0.5f };
	kb_fsine_set_fp(fs, n.op[0].osc, fd, 0.f); kb_env_set_points(fs, n.op[0].env, 2, e1);   // FM.k:45-46
	kb_fsine_set_fp(fs, n.op[1].osc, fd, 0.f); kb_env_set_points(fs, n.op[1].env, 2, e2);   // FM.k:48-49
	kb_fsine_set_fp(fs, n.op[2].osc, fc, 0.f);                                              // FM.k:51
	kb_adsr_set(fs, n.adsr, c[3].value, 0.1f, 1.f, 1.f);                                    // FM.k:53
}

// ---- Breakpoint.k / Ramp.k / Release.k (examples/Subtractive): `osc * env++ >> out` with a Fast::Sine and one Envelope
inline void kb_senv_construct(const KbFs& fs, int graph, KbSenvVoice& n) {
	kb_fsine_init(n.osc); kb_env_construct(fs, n.env);
	n.stop_when_finished = graph == KB_SY_RELEASE;                                          // Release.k:28-29; the other two never stop()
}
inline void kb_senv_on(const KbFs& fs, int graph, const KbControl* c, KbSenvVoice& n, float pitch) {
	kb_fsine_set_fp(fs, n.osc, kb_pitch_to_frequency_host(pitch), 0.f);                     // osc(pitch -> Frequency, 0)
	if (graph == KB_SY_BREAKPOINT) {                                                        // Breakpoint.k:14-16 (decay reads controls[0] too)
		const float attack = c[0].value, decay = c[0].value;
		const float pts[6] = { 0.f, 0.f, attack, 1.f, attack + decay, 0.f };
		kb_env_set_points(fs, n.env, 3, pts);
	} else if (graph == KB_SY_RAMP) {                                                       // Ramp.k:14-15
		const float pts[4] = { 0.f, 1.f, c[0].value, 0.f };
		kb_env_set_points(fs, n.env, 2, pts);
	} else {                                                                                // Release.k:14-18
		const float A = c[0].value, D = c[1].value, S = c[2].value;
		const float pts[6] = { 0.f, 0.f, A, 1.f, A + D, S };
		kb_env_set_points(fs, n.env, 3, pts);
		kb_env_set_loop(n.env, 2, 2);
	}
}

// ---- Additive/Saw.k, Additive/Square.k: on() only sets the 32 frequencies, the partials keep their phase from note to note (Saw.k:7-10, 25-27)
inline void kb_add_construct(int graph, KbAddVoice& n) {
	for (int o = 0; o < 32; o++) kb_fsine_init(n.osc[o]);
	n.square = graph == KB_SY_ADDITIVE_SQUARE ? 1 : graph == KB_SY_ADDITIVE_NYQUIST ? 2 : 0;
}
inline void kb_add_on(const KbFs& fs, KbAddVoice& n, float pitch) {
	const float f = kb_pitch_to_frequency_host(pitch);
	for (int o = 0; o < 32; o++) kb_fsine_set_f(fs, n.osc[o], f * (o + 1));
}

// ---- Modulation/AM.k:12-15, Modulation/FM.k:13-18, Modulation/FM2.k:13-19
KB_HD void kb_fsine_reset(KbFastSine& o) { o.position = 0u; o.offset = 0u; }   // Fast::Sine::reset (klang.h:5136-5140): set(frequency, 0) finds the frequency unchanged
inline void kb_smod_construct(const KbFs& fs, int graph, KbSmodVoice& n) {
	kb_fsine_init(n.carrier); kb_fsine_init(n.mod1); kb_fsine_init(n.mod2);
	kb_adsr_construct(fs, n.adsr);
	n.f0 = 0.f; n.graph = graph;
}
inline void kb_smod_on(const KbFs& fs, KbSmodVoice& n, float pitch) {
	const float f = kb_pitch_to_frequency_host(pitch);
	if (n.graph == KB_SY_AM) {                                 // the modulator keeps phase and frequency from the last note
		kb_fsine_set_fp(fs, n.carrier, f, 0.f);
		kb_adsr_set(fs, n.adsr, 0.f, .1f, 0.1f, 0.25f);
	} else {
		n.f0 = f;
		kb_fsine_reset(n.carrier); kb_fsine_reset(n.mod1);
		if (n.graph == KB_SY_MOD_FM2) kb_fsine_reset(n.mod2);
		if (n.graph == KB_SY_MOD_FM) kb_adsr_set(fs, n.adsr, 0.001f, 0.f, 1.f, 0.25f); else kb_adsr_set(fs, n.adsr, 0.5f, 0.f, 1.f, 0.25f);
	}
}

// ---- TB303 (examples/TB303.k:8-114)
inline void kb_tb_filter_reset(KbTbFilter& F) {                                            // TB303.k:25-31
	F.cutoff = 0; F.resonance = 0; F.drive = 1;
	F.z[0] = F.z[1] = F.z[2] = F.z[3] = 0;
	kb_onepole_reset(F.feedback);
}
inline void kb_tb_construct(const KbFs& fs, KbTbVoice& n) {
	memset(&n, 0, sizeof(n));
	kb_osm_construct(n.saw, 0, 0.f);
	kb_osm_construct(n.square, 1, 1.0f);
	kb_adsr_construct(fs, n.adsr);
	kb_env_construct(fs, n.env);
	n.filter.cutoff = 0; n.filter.resonance = 0; n.filter.drive = 1; n.filter.g = 1.0f;     // TB303.k:9-17
	kb_onepole_construct(n.filter.feedback, KB_OP_HPF);
}
inline void kb_tb_on(const KbFs& fs, const KbControl* c, KbTbVoice& n, float pitch) {      // TB303.k:90-97
	n.f = kb_pitch_to_frequency_host(pitch);
	kb_osm_set_fp(fs, n.square, n.f, 0.f);
	kb_osm_set_fp(fs, n.saw, n.f, 0.f);
	kb_adsr_set(fs, n.adsr, 0.00f, c[2].value, 1.f, 0.1f);
	kb_tb_filter_reset(n.filter);
	const float pts[4] = { 0.f, 1.f, c[2].value, 0.0f };
	kb_env_set_points(fs, n.env, 2, pts);
}
// Control-rate constants of the TB303 per-sample code, evaluated on the host with the host libm exactly as the
// reference would evaluate them every sample (they depend on controls only): TB303.k:40 (r), :34 (feedback HPF
// coefficients via OnePole::HPF::init, klang.h:5535-5541), :67 (soft-clip denominator, double).
struct KbTbBlock { float r; float hpf_f, hpf_b0, hpf_b1, hpf_a1; double clip_den; float resonance, drive, c0sq_nyq; int is_square; };
inline KbTbBlock kb_tb_block(const KbFs& fs, const KbControl* c) {
	KbTbBlock k;
	k.is_square = c[3].value == 1;
	k.resonance = (k.is_square ? 0.9f : 1.f) * c[1].value;
	k.drive = c[4].value;
	k.r = (1.f - ::expf(-0.03f * k.resonance)) / (1.f - ::expf(-3.f));
	k.hpf_f = 10.f + 490.f * c[1].value;
	const float e = ::expf(-k.hpf_f * fs.w);
	k.hpf_b0 = 0.5f * (1.f + e); k.hpf_b1 = -k.hpf_b0; k.hpf_a1 = e;
	k.clip_den = 0.1 * k.drive + ::tanhf(k.drive);
	k.c0sq_nyq = (c[0].value * c[0].value) * fs.nyquist;
	return k;
}

// ---- SynTHX (examples/SynTHX.k:8-183)
static const float kb_sx_transposition[26] = { 0.00f,0.0625f, 0.08f,0.078125f, 0.17f,0.09375f, 0.25f,0.125f, 0.33f,0.15625f, 0.41f,0.1875f,
                                               0.50f,0.25f, 0.58f,0.3125f, 0.66f,0.375f, 0.75f,0.5f, 0.83f,0.625f, 0.91f,0.75f, 1.00f,1.f };   // SynTHX.k:14-28
static const float kb_sx_detunes[12] = { 0.f,0.0f, 0.6f,0.015f, 0.7f,0.05f, 0.8f,0.1f, 0.9f,0.5f, 1.0f,2.00f };                              // SynTHX.k:30-33

// Partial::set(transpose, detune): runs on the host inside on() and on the device once per block     SynTHX.k:37-45
KB_HD void kb_sx_partial_set2(const KbFs& fs, KbSxPartial& p, float tr_at, float dt_at) {
	const float detuned = tr_at * (1.f + p.seed * dt_at);
	float f = p.f0 + detuned * p.range;
	if (f < 0) f = -f;
	else if (f >= fs.nyquist) f = f * 0.5f;
	kb_osm_set_f(fs, p.osc, f);
}

inline float kb_sx_tr_at(float transpose) { float px[13], py[13]; for (int i = 0; i < 13; i++) { px[i] = kb_sx_transposition[2 * i]; py[i] = kb_sx_transposition[2 * i + 1]; } return kb_env_at(px, py, 13, transpose); }
inline float kb_sx_dt_at(float detune) { float px[6], py[6]; for (int i = 0; i < 6; i++) { px[i] = kb_sx_detunes[2 * i]; py[i] = kb_sx_detunes[2 * i + 1]; } return kb_env_at(px, py, 6, detune); }
inline void kb_sx_construct(const KbFs& fs, KbSxVoice& n) {
	memset(&n, 0, sizeof(n));
	for (int k = 0; k < 11; k++) {
		n.notes[k].frequency = 1000.f;                                                     // Oscillator::frequency  klang.h:2855
		for (int p = 0; p < 4; p++) for (int q = 0; q < 3; q++) kb_osm_construct(n.notes[k].partial[p][q].osc, 0, 0.f);
	}
	kb_adsr_construct(fs, n.adsr);
}
// Partial::set(f, transpose, detune)                                          SynTHX.k:48-57
inline void kb_sx_partial_set3(const KbFs& fs, KbSxPartial& p, float f, float tr_at, float dt_at) {
	p.f0 = 0 * f;
	p.range = 1 * f;
	p.right = kb_randomf(0.f, 1.f) > 0.5;
	p.seed = kb_randomf(-1.f, 1.f);
	if (fabsf(p.seed) < 0.0003) p.seed = kb_randomf(0.f, 1.f) > 0.5 ? (float)0.0003 : (float)-0.0003;
	kb_sx_partial_set2(fs, p, tr_at, dt_at);
}
// MyNote::on: `controls[3] ? minor : major` converts through `operator Control*()` (never null) => minor   SynTHX.k:154-161
inline void kb_sx_on(const KbFs& fs, const KbControl* c, KbSxVoice& n, float pitch) {
	static const float minor[11] = { 26, 33, 38, 45, 50, 53, 57, 64, 67, 71, 76 };        // SynTHX.k:134-140
	const float tr_at = kb_sx_tr_at(c[2].value), dt_at = kb_sx_dt_at(c[1].value);
	for (int k = 0; k < 11; k++) {
		for (int rep = 0; rep < 2; rep++) {                                                // the reference calls set() twice
			const float f0 = kb_pitch_to_frequency_host(minor[k] - 26 + pitch);
			n.notes[k].frequency = f0;                                                     // Additive::set  SynTHX.k:83-91
			for (int p = 0; p < 4; p++) {
				const float f = f0 * (1 + p);
				for (int q = 0; q < 3; q++) kb_sx_partial_set3(fs, n.notes[k].partial[p][q], f, tr_at, dt_at);
			}
		}
	}
	kb_adsr_set(fs, n.adsr, c[0].value, 0.f, 1.f, 2.0f);
}

// ---- effects: constructors and prepare()
inline void kb_pingpong_construct(KbFxHdr& h, KbPingPong& p, long long ring0) {          // PingPong.k:6-20
	memset(&p, 0, sizeof(p));
	kb_delay_construct(p.left, 192000, ring0); kb_delay_construct(p.right, 192000, ring0 + 192000 + 1);
	kb_bosc_init(p.lfo);
	kb_biquad_construct(p.dc[0], KB_BQ_HPF); kb_biquad_construct(p.dc[1], KB_BQ_HPF);
	h.controls[0] = kb_dial(0.0f, 0.999f, 0.5f); h.controls[1] = kb_dial(0.001f, 1.0f, 0.5f);
	h.controls[2] = kb_dial(0.0f, 1.0f, 0.0f);   h.controls[3] = kb_dial(0.01f, 1.0f, 0.0f);
	h.controls[4] = kb_dial(0.001f, 2.0f, 1.0f); h.controls[5] = kb_dial(0.f, 1.f, 0.f);
}
inline void kb_pingpong_prepare(const KbFs& fs, KbPingPong& p) {                           // PingPong.k:36-40
	kb_biquad_set(fs, p.dc[0], 50.f, 1.f);
	kb_biquad_set(fs, p.dc[1], 50.f, 1.f);
}

inline void kb_reverb_construct(KbFxHdr& h, KbReverb& rv, long long ring0) {               // Reverb.k:12, 99-111, 119-123
	memset(&rv, 0, sizeof(rv));
	long long r = ring0;
	// (every ring starts on a 16-byte boundary and is padded to a multiple of 4 floats: the ring windows travel as 1-D bulk copies)
	kb_delay_construct(rv.dl, 21600, r); r += 21604;
	kb_delay_construct(rv.dr, 21600, r); r += 21604;
	for (int c = 0; c < 2; c++) { kb_biquad_construct(rv.lpf[c], KB_BQ_LPF); kb_biquad_construct(rv.hpf[c], KB_BQ_HPF); }
	for (int k = 0; k < 2; k++) for (int i = 0; i < 4; i++) {
		kb_delay_construct(rv.mid[k].d[i].delay, 192000, r); r += 192004; kb_biquad_construct(rv.mid[k].d[i].filter, KB_BQ_LPF);
		kb_delay_construct(rv.late[k].d[i].delay, 192000, r); r += 192004; kb_biquad_construct(rv.late[k].d[i].filter, KB_BQ_LPF);
	}
	h.controls[0] = kb_dial(0.f, 1.f, 0.f); h.controls[1] = kb_dial(0.f, 1.f, 1.f); h.controls[2] = kb_dial(0.f, 1.f, 0.f);
	h.controls[3] = kb_dial(0.f, 1.f, 0.f); h.controls[4] = kb_dial(0.f, 1.f, 1.f); h.controls[5] = kb_dial(0.f, 100.f, 10.f);
	h.controls[6] = kb_dial(0.f, 1.f, 1.f); h.controls[7] = kb_dial(0.01f, 1.f, 1.f); h.controls[8] = kb_dial(0.01f, 1.f, 1.f);
	h.controls[9] = kb_dial(0.f, 0.2f, 0.f);
}
#define KB_REVERB_RING_FLOATS (2LL * 21604 + 16LL * 192004)
#define KB_PINGPONG_RING_FLOATS (2LL * 192001)
// EarlyReflections::update                                                   Reverb.k:23-52
inline void kb_rv_early_update(const KbFs& fs, KbReverb& rv) {
	static const float primes[20] = { 2,3,5,7,11, 13,17,19,23,29, 31,37,41,43,47, 53,59,61,67,71 };
	rv.count = 10 + (int)(rv.size * (float)10.999);
	const float scale = 50.f / primes[rv.count - 1];
	const float ms = fs.f / 1000.f;
	for (int r = 0; r < rv.count; r++) {
		rv.times[r] = ((50.f + primes[r] * scale) * ms * kb_randomf(0.9f, 1.1f));
		const float x = (float)(r + 1.f) / rv.count;
		const float g = kb_randomf(0.5f, 1.5f) * ::expf(-3.f * x);
		const float pan = kb_randomf(0.f, 1.f);
		rv.gl[r] = g * (1.f - pan);
		rv.gr[r] = g * pan;
	}
}
// LateReflections::set + FilteredDelay::set                                  Reverb.k:124-143
inline void kb_rv_late_set(const KbFs& fs, KbRvLate& L, const float* delays, float dampening, float gain) {
	for (int i = 0; i < 4; i++) {
		const float time = delays[i] * kb_randomf(.9f, 1.1f);
		kb_delay_set(L.d[i].delay, time * fs.f / 1000.f);
		kb_biquad_set_f(fs, L.d[i].filter, dampening);
		L.d[i].gain = gain;
	}
}
// Reverb::prepare -> Reflections::set -> EarlyReflections::set               Reverb.k:238-242, 181-204, 63-73
inline void kb_reverb_prepare(const KbFs& fs, KbFxHdr& h, KbReverb& rv) {
	bool changed = false;                                                                   // Controls::changed  klang.h:1914-1923
	for (int c = 0; c < 10; c++) if (h.controls[c].value != h.cached[c]) { h.cached[c] = h.controls[c].value; changed = true; }
	if (!changed) return;
	srand(272839);
	float length = (h.controls[5].value / 10.f) * 1000.f + 50.f;
	const float size = h.controls[6].value;
	float dampening1 = h.controls[7].value, dampening2 = h.controls[8].value;
	length *= 1 / 1000.f;
	if (rv.length != length || rv.size != size) {
		rv.length = length; rv.size = size;
		kb_rv_early_update(fs, rv);
		for (int c = 0; c < 2; c++) kb_biquad_set_f(fs, rv.hpf[c], 100.f);
		for (int c = 0; c < 2; c++) kb_biquad_set_f(fs, rv.lpf[c], 15000.f);
	}
	dampening1 *= 10000.f;
	dampening2 *= dampening1;
	static const float delays1[4] = { 7, 11, 13, 17 };
	static const float delays2[4] = { 19, 23, 29, 31 };
	kb_rv_late_set(fs, rv.mid[0], delays1, dampening1, (float)0.25);
	kb_rv_late_set(fs, rv.mid[1], delays1, dampening1, (float)0.25);
	kb_rv_late_set(fs, rv.late[0], delays2, dampening2, (float)0.35);
	kb_rv_late_set(fs, rv.late[1], delays2, dampening2, (float)0.35);
}

inline void kb_dpingpong_construct(KbFxHdr& h, KbDPingPong& p, long long ring0) {          // Delay/PingPong.k:11-20
	memset(&p, 0, sizeof(p));
	kb_delay_construct(p.l, 192000, ring0); kb_delay_construct(p.r, 192000, ring0 + 192001);
	h.controls[0] = kb_dial(0.f, 1.f, 0.25f); h.controls[1] = kb_dial(0.f, 1.f, 0.5f);
	h.controls[2] = kb_dial(0.f, 1.f, 0.5f);  h.controls[3] = kb_dial(0.f, 1.f, 0.5f);
}
inline void kb_dreverb_construct(KbFxHdr& h, KbDReverb& p, long long ring0) {              // Delay/Reverb.k:51-55
	memset(&p, 0, sizeof(p));
	kb_delay_construct(p.feedforward, 192000, ring0); kb_delay_construct(p.feedback, 192000, ring0 + 192001);
	kb_biquad_construct(p.filter, KB_BQ_LPF);
	h.controls[0] = kb_dial(0.f, 0.5f, 0.4f); h.controls[1] = kb_dial(0.f, 0.4f, 0.1f); h.controls[2] = kb_dial(500.f, 5000.f, 1500.f);
}

// ---- Delay/Echo.k:16-23 and Delay/Feedback.k:16-24: one frame (host + device: tests/host/onedelay_check.cpp renders them with g++)
#define KB_ONEDELAY_RING_FLOATS 192008LL
KB_D float kb_echo_frame(const KbFs& fs, const KbFxHdr& h, KbOneDelayFx& s, float* rings, float in) {
	float* ring = rings + s.delay.ring;
	const float time = h.controls[0].value * fs.f, gain = h.controls[1].value;
	kb_delay_write(s.delay, ring, in);                                        // in >> delay
	return in + kb_delay_tap_f(s.delay, ring, time) * gain;                  // in + delay(time) * gain >> out
}
KB_D float kb_feedback_frame(const KbFs& fs, const KbFxHdr& h, KbOneDelayFx& s, float* rings, float in) {
	float* ring = rings + s.delay.ring;
	const float time = h.controls[0].value * fs.f, gain = h.controls[1].value;
	const float out = in + kb_delay_tap_f(s.delay, ring, time) * gain;       // in + delay(time) * gain >> out
	kb_delay_write(s.delay, ring, out);                                       // delay << out
	return out;
}

// Filtering/IIR.k:14-26: f = cube(controls[0]); filtered = (1 - f) * last + f * dry; last = filtered   (a true fp32 recurrence: one lane)
KB_HD float kb_iir_frame(const KbFxHdr& h, KbIirFx& s, float in) {
	const float x = h.controls[0].value, f = x * x * x;
	const float filtered = (1 - f) * s.last + f * in;
	s.last = filtered;
	return filtered;
}

// Filtering/WahWah.k:18-27: mod = lfo(rate) * 0.5 + 0.5f; in >> lpf(sqr(mod) * f, Q) >> out.  Biquad::Filter::set runs every sample
// (cosf / sinf of the moving cutoff, klang.h:5584-5600); the filter state is the serial chain: one lane per instance.
KB_HD float kb_wahwah_frame(const KbFs& fs, const KbFxHdr& h, KbWahWahFx& s, float in) {
	const float f = h.controls[0].value, Q = h.controls[1].value, rate = h.controls[2].value;
	kb_fsine_set_f(fs, s.lfo, rate);
	const float mod = kb_fsine_tick(s.lfo) * 0.5f + 0.5f;
	kb_biquad_set(fs, s.lpf, (mod * mod) * f, Q);
	return kb_biquad_tick(s.lpf, in);
}

// Modulation/Flanger.k:17-24, ModDelay.k:18-26, Chorus.k:17-28: the input is written to the line, which is tapped at LFO-modulated times
// (feed-forward: the two-sweep schedule of Echo.k applies; frame-sequential for now).  ModDelay.k smooths its depth control per
// sample (Control::smooth, klang.h:1715-1716), which lives in the header's control block.
KB_D float kb_moddelay_frame(int graph, const KbFs& fs, KbFxHdr& h, KbModDelayFx& s, float* rings, float in) {
	float* ring = rings + s.delay.ring;
	if (graph == KB_FX_FLANGER) {
		const float rate = h.controls[0].value, depth = h.controls[1].value / 1000.f;
		kb_osm_set_f(fs, s.tri, rate);
		const float mod = kb_osm_tick(s.tri) * depth + depth;
		kb_delay_write(s.delay, ring, in);
		return in + kb_delay_tap_f(s.delay, ring, mod * fs.f);
	}
	if (graph == KB_FX_MODDELAY) {
		const float rate = h.controls[0].value;
		const float sm = kb_control_smooth(h.controls[1]);
		const float depth = (sm * sm * sm) / 10.f;
		kb_fsine_set_f(fs, s.lfo[0], rate);
		const float mod = kb_fsine_tick(s.lfo[0]) * depth + depth;
		kb_delay_write(s.delay, ring, in);
		return kb_delay_tap_f(s.delay, ring, mod * fs.f);
	}
	const float rates[3] = { 2.5f, 3.f, 3.5f }, depths[3] = { 0.45f, 0.5f, 0.55f };
	kb_delay_write(s.delay, ring, in);
	float acc = in;
	for (int k = 0; k < 3; k++) {
		kb_fsine_set_f(fs, s.lfo[k], rates[k]);
		const float t = (kb_fsine_tick(s.lfo[k]) * depths[k] + depths[k]) * fs.f / 1000.f;
		acc = acc + kb_delay_tap_f(s.delay, ring, t);
	}
	return 0.5f * acc;
}

// Flanger.k / Modulation/Chorus.k, time-parallel (feed-forward taps, LFOs in closed form): a write sweep that also STASHES the value each
// slot held before (old[t] = what frame t overwrote), then a read sweep in which frame t sees every slot as the frame-sequential run
// would: a slot that a LATER frame of this block overwrote is read from the stash, every other slot from the ring.  Exact for any
// delay (a zero delay included: its interpolation partner is the next frame's slot) as long as n < SIZE.
KB_D float kb_line_slot(const float* ring, const float* old, int p0, int n, int SIZE, int t, int k) {
	if (k < SIZE) {                                                           // (slot SIZE is the guard float: never written)
		int writer = k - p0;                                                  // the frame of this block that writes slot k, if < n
		if (writer < 0) writer += SIZE;
		if (writer > t && writer < n) return old[writer];
	}
	return ring[k];
}
KB_D float kb_line_tap_f(const float* ring, const float* old, int p0, int n, int SIZE, int t, float delay) {     // Delay::tap(float) as frame t finds the line
	const int position = (p0 + t + 1) % SIZE;
	float read = (float)(position - 1) - delay;
	if (read < 0.f) read += SIZE;
	const int i = (int)read;
	const float fraction = read - i;
	const int j = (i + 1) % SIZE;
	const float a = kb_line_slot(ring, old, p0, n, SIZE, t, i), b = kb_line_slot(ring, old, p0, n, SIZE, t, j);
	return a + fraction * (b - a);
}
// ModDelay.k joins them with one serial pre-pass: its depth is cube(controls[1].smooth()) / 10 (ModDelay.k:20), and Control::smooth is a
// one-pole recurrence per sample — kb_modline_begin runs it for the block's n frames on one lane and leaves the depth of every frame
// in `depth`; everything else is the same two sweeps.
KB_HD void kb_modline_begin(int graph, const KbFs& fs, KbFxHdr& h, KbModDelayFx& s, float* depth, int n) {   // what the block's first frame does to the LFO settings
	if (graph == KB_FX_FLANGER) kb_osm_set_f(fs, s.tri, h.controls[0].value);
	else if (graph == KB_FX_MODDELAY) {
		kb_fsine_set_f(fs, s.lfo[0], h.controls[0].value);
		for (int t = 0; t < n; t++) { const float sm = kb_control_smooth(h.controls[1]); depth[t] = (sm * sm * sm) / 10.f; }
	} else { const float rates[3] = { 2.5f, 3.f, 3.5f }; for (int k = 0; k < 3; k++) kb_fsine_set_f(fs, s.lfo[k], rates[k]); }
}
KB_D void kb_modline_write_at(const KbDelay& d, float* rings, float* old, int t, float in) {
	float* slot = rings + d.ring + (d.position + t) % d.SIZE;
	old[t] = *slot;
	*slot = in;
}
KB_D float kb_modline_read_at(int graph, const KbFs& fs, const KbFxHdr& h, const KbModDelayFx& s, const float* rings, const float* old, const float* depth_row,
                              int n, int t, float in) {
	const float* ring = rings + s.delay.ring;
	if (graph == KB_FX_MODDELAY) {
		const float depth = depth_row[t];
		const float mod = kb_fsine_value(s.lfo[0].position + (uint32_t)t * (uint32_t)s.lfo[0].increment + s.lfo[0].offset) * depth + depth;
		return kb_line_tap_f(ring, old, s.delay.position, n, s.delay.SIZE, t, mod * fs.f);
	}
	if (graph == KB_FX_FLANGER) {
		const float depth = h.controls[1].value / 1000.f;
		const float mod = kb_osm_at(s.tri, (uint32_t)t) * depth + depth;
		return in + kb_line_tap_f(ring, old, s.delay.position, n, s.delay.SIZE, t, mod * fs.f);
	}
	const float depths[3] = { 0.45f, 0.5f, 0.55f };
	float acc = in;
	for (int k = 0; k < 3; k++) {
		const float sine = kb_fsine_value(s.lfo[k].position + (uint32_t)t * (uint32_t)s.lfo[k].increment + s.lfo[k].offset);
		acc = acc + kb_line_tap_f(ring, old, s.delay.position, n, s.delay.SIZE, t, (sine * depths[k] + depths[k]) * fs.f / 1000.f);
	}
	return 0.5f * acc;
}
KB_HD void kb_modline_end(int graph, KbModDelayFx& s, int n) {
	if (graph == KB_FX_FLANGER) kb_osm_advance(s.tri, (uint32_t)n);
	else for (int k = 0; k < (graph == KB_FX_MODDELAY ? 1 : 3); k++) s.lfo[k].position += (uint32_t)n * (uint32_t)s.lfo[k].increment;
	s.delay.position = (s.delay.position + n) % s.delay.SIZE;
}

// Echo.k, time-parallel: the line is only ever fed the INPUT, so a block is two independent sweeps — write all n inputs into the ring, then
// every output sample from its own tap.  Exact as long as no tap of the block reads a slot that a LATER sample of the same block
// overwrites: n + delay + 2 < SIZE at the far end, and delay >= 1 at the near end (with a delay below one frame the interpolation's
// second slot is the NEXT frame's — weight 0, but an exact zero times a different value can flip the sign of a zero).
// kb_echo_parallel_ok checks both; frame t sees the line with position (p0 + t + 1) mod SIZE.
KB_HD bool kb_echo_parallel_ok(const KbFs& fs, int n, float control0) {
	const float delay = control0 * fs.f;
	return delay >= 1.f && (double)n + (double)delay + 3.0 < 192000.0;
}
KB_D void kb_echo_write_at(const KbOneDelayFx& s, float* rings, int t, float in) { rings[s.delay.ring + (s.delay.position + t) % s.delay.SIZE] = in; }
KB_D float kb_echo_read_at(const KbFs& fs, const KbFxHdr& h, const KbOneDelayFx& s, const float* rings, int t, float in) {
	KbDelay d = s.delay;
	d.position = (s.delay.position + t + 1) % s.delay.SIZE;
	return in + kb_delay_tap_f(d, rings + s.delay.ring, h.controls[0].value * fs.f) * h.controls[1].value;
}

// Feedback.k, chunk-parallel: frame t reads the line `delay` frames back and the line is fed the OUTPUT, so frames closer together than the
// delay do not see each other.  A block is cut into chunks of kb_feedback_chunk frames (<= floor(delay) - 1, so both interpolation
// slots of every frame in a chunk were written before the chunk began); chunks run in order, the frames of a chunk in any order
// (kb_feedback_at).  Chunk 0 = not parallelisable (delay below 3 frames, or the far end n + delay + 2 >= SIZE): frame-sequential.
KB_HD int kb_feedback_chunk(const KbFs& fs, int n, float control0) {
	const float delay = control0 * fs.f;
	if (!(delay >= 3.f) || !((double)n + (double)delay + 3.0 < 192000.0)) return 0;
	return (int)delay - 1;
}
KB_D float kb_feedback_at(const KbFs& fs, const KbFxHdr& h, const KbOneDelayFx& s, float* rings, int t, float in) {
	float* ring = rings + s.delay.ring;
	KbDelay d = s.delay;
	d.position = (s.delay.position + t) % s.delay.SIZE;                       // the line as frame t finds it: t frames written since the block began
	const float out = in + kb_delay_tap_f(d, ring, h.controls[0].value * fs.f) * h.controls[1].value;
	ring[d.position] = out;                                                   // delay << out
	return out;
}

// ---- FM.k per-sample half (host + device: tests/host/fm_check.cpp renders it with g++)
// Operator::process (klang.h:4163-4167): OSCILLATOR::set(+in) -> Fast::Sine::set(relative): offset = phase * twoPi through
// Fast::Phase::operator=(klang::Phase) (klang.h:5160-5162, 4993-4997, Q2); Sine::process; out *= env++ * amp
KB_HD float kb_fm_op_tick(const KbFs& fs, KbFmOp& o) {
	o.osc.offset = kb_phase_from_radians(o.in * KB_TWO_PI_F);
	float out = kb_fsine_tick(o.osc);
	out *= kb_env_tick(fs, o.env) * o.amp;
	return out;
}
// FM.k:61-73: op1 * I1 >> op2 * I2 >> op3 >> out; out *= adsr++ * 0.1f
KB_HD float kb_fm_tick(const KbFs& fs, float i1, float i2, KbFmVoice& n, int& note_stage) {
	n.op[0].amp = i1;
	n.op[1].amp = i2;
	n.op[1].in = kb_fm_op_tick(fs, n.op[0]);
	n.op[2].in = kb_fm_op_tick(fs, n.op[1]);
	float out = kb_fm_op_tick(fs, n.op[2]);
	out *= kb_env_tick(fs, n.adsr) * 0.1f;
	if (n.adsr.stage == KB_ENV_OFF) note_stage = KB_NOTE_OFF;
	return out;
}

// Breakpoint.k:20 / Ramp.k:19 / Release.k:27-29: osc * env++ >> out; Release.k stops the note when the envelope has finished
KB_HD float kb_senv_tick(const KbFs& fs, KbSenvVoice& n, int& note_stage) {
	const float o = kb_fsine_tick(n.osc);
	const float out = o * kb_env_tick(fs, n.env);
	if (n.stop_when_finished && n.env.stage == KB_ENV_OFF) note_stage = KB_NOTE_OFF;
	return out;
}

// ---- elementwise effects: Gain/Pan.k:16-21, Gain/RM.k:17-24, Gain/Tremolo.k:22-29, Distortion/Clipping.k:14-25, Functions.k, Mute.k.  Sample `t` of a block as a
// pure function of the input sample: the LFO of RM.k / Tremolo.k is a Fast::Sine on an integer phase ramp, and `lfo(rate)` sets a
// frequency that is constant over the block (controls move between blocks), so Sine::set runs once per block on the host.
// c0, c1 = controls[0], controls[1]; channel = 0 left / mono, 1 right.
KB_HD float kb_ew_sample(int graph, float c0, float c1, const KbFastSine& lfo, int channel, uint32_t t, float in) {
	if (graph == KB_FX_PAN) return channel == 0 ? in * (1 - c0) : in * c0;
	if (graph == KB_FX_MUTE) return in * (c0 ? 0.f : 1.f);                    // Mute.k:20-31 (the product keeps the sign of a zero)
	if (graph == KB_FX_CLIPPING || graph == KB_FX_FUNCTIONS) {               // Clipping.k:14-25; Functions.k:4-11, 25-29 hardclip(in * gain)
		in *= c0;
		if (in > 1) in = 1; else if (in < -1) in = -1;
		return in;
	}
	const float s = kb_fsine_value(lfo.position + t * (uint32_t)lfo.increment + lfo.offset);
	const float mod = graph == KB_FX_RM ? s : s * c1 + (1 - c1);
	return in * mod;
}

// Modulation/AM.k:22-30, FM.k:25-32, FM2.k:26-36: `modulator(rate)` / `carrier(f0 + mod)` = Fast::Sine::set(f) (recomputes the integer
// increment whenever f differs from the cached frequency) followed by one tick; c0..c2 = controls[0..2]
KB_HD float kb_smod_tick(const KbFs& fs, float c0, float c1, float c2, KbSmodVoice& n, int& note_stage) {
	float out;
	if (n.graph == KB_SY_AM) {
		const float mod_rate = c0 * n.carrier.frequency, mod_depth = c1;
		kb_fsine_set_f(fs, n.mod1, mod_rate);
		const float mod = kb_fsine_tick(n.mod1) * mod_depth + (1 - mod_depth);
		out = kb_fsine_tick(n.carrier) * mod;
		out = out * kb_env_tick(fs, n.adsr);
	} else if (n.graph == KB_SY_MOD_FM) {
		const float mod_rate = c0 * n.f0, mod_depth = c1 * n.f0;
		kb_fsine_set_f(fs, n.mod1, mod_rate);
		const float mod = kb_fsine_tick(n.mod1) * mod_depth;
		kb_fsine_set_f(fs, n.carrier, n.f0 + mod);
		out = kb_fsine_tick(n.carrier) * kb_env_tick(fs, n.adsr);
	} else {
		const float mod_rate = c0 * n.f0, d1 = c1 * n.f0, d2 = c2 * n.f0;
		kb_fsine_set_f(fs, n.mod1, mod_rate);
		const float mod1 = kb_fsine_tick(n.mod1) * d1;
		kb_fsine_set_f(fs, n.mod2, n.f0 + mod1);
		const float mod2 = kb_fsine_tick(n.mod2) * d2;
		kb_fsine_set_f(fs, n.carrier, n.f0 + mod2);
		out = kb_fsine_tick(n.carrier) * kb_env_tick(fs, n.adsr);
	}
	if (n.adsr.stage == KB_ENV_OFF) note_stage = KB_NOTE_OFF;
	return out;
}

// Time-parallel forms of the voices whose only recurrence is ONE envelope (Breakpoint.k / Ramp.k / Release.k, and Modulation/AM.k, whose
// modulator frequency c0 * carrier.frequency is constant over a block): kb_es_begin = what the first tick of the block does to the
// oscillator settings, kb_es_at = sample `t` from the block-start state and the envelope value e of that sample, kb_es_end = the
// oscillators after `ticks` ticks.  Same contract as kb_fm_at; kb_esine_tiled_kernel (kb_tiled.cuh) runs them.
KB_HD KbEnv& kb_es_env(KbSenvVoice& n) { return n.env; }
KB_HD KbEnv& kb_es_env(KbSmodVoice& n) { return n.adsr; }
KB_HD bool kb_es_stops(const KbSenvVoice& n) { return n.stop_when_finished != 0; }
KB_HD bool kb_es_stops(const KbSmodVoice&) { return true; }
KB_HD void kb_es_begin(const KbFs&, KbSenvVoice&, float) {}
KB_HD void kb_es_begin(const KbFs& fs, KbSmodVoice& n, float c0) { kb_fsine_set_f(fs, n.mod1, c0 * n.carrier.frequency); }   // AM.k:23, 26
KB_HD float kb_es_at(const KbSenvVoice& n, uint32_t t, float, float e) {
	return kb_fsine_value(n.osc.position + t * (uint32_t)n.osc.increment + n.osc.offset) * e;
}
KB_HD float kb_es_at(const KbSmodVoice& n, uint32_t t, float c1, float e) {
	const float mod = kb_fsine_value(n.mod1.position + t * (uint32_t)n.mod1.increment + n.mod1.offset) * c1 + (1 - c1);
	float out = kb_fsine_value(n.carrier.position + t * (uint32_t)n.carrier.increment + n.carrier.offset) * mod;
	out = out * e;
	return out;
}
KB_HD void kb_es_end(KbSenvVoice& n, uint32_t ticks) { n.osc.position += ticks * (uint32_t)n.osc.increment; }
KB_HD void kb_es_end(KbSmodVoice& n, uint32_t ticks) {
	n.carrier.position += ticks * (uint32_t)n.carrier.increment;
	n.mod1.position += ticks * (uint32_t)n.mod1.increment;
}
// the oscillator fields a block changes (the envelope is written back by its own lane)
KB_HD void kb_es_writeback(KbSenvVoice& dst, const KbSenvVoice& m) { dst.osc = m.osc; }
KB_HD void kb_es_writeback(KbSmodVoice& dst, const KbSmodVoice& m) { dst.carrier = m.carrier; dst.mod1 = m.mod1; }

// Additive/Saw.k:12-16, Square.k:12-19: out = 0; out += osc[o] / (o + 1) in partial order; Square.k ticks only the odd harmonics whose
// frequency lies below Nyquist.  kb_add_tick is the per-tick form, kb_add_at sample `t` of a block as a pure function of the block-start
// phases (no value is carried from sample to sample: the voice is 32 integer phase ramps); kb_add_block_end leaves the voice as
// `ticks` calls of kb_add_tick would.
KB_HD bool kb_add_partial_on(const KbFs& fs, const KbAddVoice& n, int o) {
	if (n.square == 0) return true;
	return (n.square == 2 || ((o + 1) % 2)) && n.osc[o].frequency < fs.nyquist;      // Square.k:15-17, Nyquist.k:14-16
}
KB_HD float kb_add_tick(const KbFs& fs, KbAddVoice& n) {
	float out = 0;
	for (int o = 0; o < 32; o++) if (kb_add_partial_on(fs, n, o)) out += kb_fsine_tick(n.osc[o]) / (o + 1);
	return out;
}
KB_HD float kb_add_at(const KbFs& fs, const KbAddVoice& n, uint32_t t) {
	float out = 0;
	for (int o = 0; o < 32; o++)
		if (kb_add_partial_on(fs, n, o)) out += kb_fsine_value(n.osc[o].position + t * (uint32_t)n.osc[o].increment + n.osc[o].offset) / (o + 1);
	return out;
}
KB_HD void kb_add_block_end(const KbFs& fs, KbAddVoice& n, uint32_t ticks) {
	for (int o = 0; o < 32; o++) if (kb_add_partial_on(fs, n, o)) n.osc[o].position += ticks * (uint32_t)n.osc[o].increment;
}

// The same sample as a pure function of the sample index: tick `t` of a block whose first tick finds the voice in state `n`.
// The three operator phases are integer ramps and an operator's phase offset is the previous operator's output of the SAME
// sample, so the only values carried from sample to sample are the four envelopes (e0..e2 the operators', ea the ADSR), which
// the caller supplies for tick t.  kb_fm_block_end leaves the voice as `ticks` calls of kb_fm_tick would (envelopes aside).
struct KbFmSample { float y0, y1, out; uint32_t off0, off1, off2; };
KB_HD KbFmSample kb_fm_at(const KbFmVoice& n, uint32_t t, float i1, float i2, float e0, float e1, float e2, float ea) {
	KbFmSample r;
	r.off0 = kb_phase_from_radians(n.op[0].in * KB_TWO_PI_F);
	r.y0 = kb_fsine_value(n.op[0].osc.position + t * (uint32_t)n.op[0].osc.increment + r.off0);
	r.y0 *= e0 * i1;
	r.off1 = kb_phase_from_radians(r.y0 * KB_TWO_PI_F);
	r.y1 = kb_fsine_value(n.op[1].osc.position + t * (uint32_t)n.op[1].osc.increment + r.off1);
	r.y1 *= e1 * i2;
	r.off2 = kb_phase_from_radians(r.y1 * KB_TWO_PI_F);
	float out = kb_fsine_value(n.op[2].osc.position + t * (uint32_t)n.op[2].osc.increment + r.off2);
	out *= e2 * n.op[2].amp;
	out *= ea * 0.1f;
	r.out = out;
	return r;
}
KB_HD void kb_fm_block_end(KbFmVoice& n, uint32_t ticks, float i1, float i2, const KbFmSample& last) {
	if (ticks == 0) return;
	n.op[0].amp = i1; n.op[1].amp = i2;
	n.op[1].in = last.y0; n.op[2].in = last.y1;
	n.op[0].osc.offset = last.off0; n.op[1].osc.offset = last.off1; n.op[2].osc.offset = last.off2;
	for (int k = 0; k < 3; k++) n.op[k].osc.position += ticks * (uint32_t)n.op[k].osc.increment;
}

// =========================================================================================== DEVICE halves
#ifdef __CUDACC__

// Filter.k:29-36 inside Note::process(buffer) (klang.h:4295-4303)
KB_D float kb_sub_tick(const KbFs& fs, KbSubVoice& n, int& note_stage) {
	kb_biquad_set(fs, n.filter, kb_env_tick(fs, n.env), 10.f);
	float out = kb_biquad_tick(n.filter, kb_osm_tick(n.osc));
	out *= kb_env_tick(fs, n.adsr);
	if (n.adsr.stage == KB_ENV_OFF) note_stage = KB_NOTE_OFF;          // stop()  klang.h:4277-4280
	return out;
}
// SuperSaw.k:25-33
KB_D float kb_ssaw_tick(const KbFs& fs, KbSsawVoice& n, int& note_stage) {
	float out = 0;
	for (int k = 0; k < 7; k++) out += kb_osm_tick(n.osc[k]) / 7;
	out *= kb_env_tick(fs, n.adsr);
	if (n.adsr.stage == KB_ENV_OFF) note_stage = KB_NOTE_OFF;
	return out;
}
// TB303.k:37-55 (Filter::set), :57-79 (shape, clips, Filter::process), :103-113 (MyNote::process)
KB_D float kb_tb_tick(const KbFs& fs, const KbTbBlock& B, KbTbVoice& n, int& note_stage) {
	const float osc = B.is_square ? (kb_osm_tick(n.square) * 0.5f) : kb_osm_tick(n.saw);
	const float e = kb_env_tick(fs, n.env);
	float cutoff = (n.f + B.c0sq_nyq) * (e * e);
	KbTbFilter& F = n.filter;
	if (cutoff > fs.nyquist) cutoff = fs.nyquist;
	if (F.cutoff != cutoff || F.resonance != B.resonance || F.drive != B.drive) {
		F.cutoff = cutoff; F.resonance = B.resonance; F.drive = B.drive;
		F.r = B.r;
		const float fx = cutoff * fs.inv * KB_ROOT2_INV_F;
		F.b0 = (0.00045522346f + 6.1922189f * fx) / (1.f + 12.358354f * fx + 4.4156345f * (fx * fx));
		float k = fx*(fx*(fx*(fx*(fx*(fx+7198.6997f)-5837.7917f)-476.47308f)+614.95611f)+213.87126f)+16.998792f;
		float g = k * 0.058823529411764705882352941176471f;
		g = (g - 1.f) * F.r + 1.f;
		g = (g * (1.f + F.r));
		k = k * F.r;
		F.k = k; F.g = g;
	}
	if (F.feedback.f != B.hpf_f) { F.feedback.f = B.hpf_f; F.feedback.b0 = B.hpf_b0; F.feedback.b1 = B.hpf_b1; F.feedback.a1 = B.hpf_a1; }   // setHPF  TB303.k:33-35
	const float a = kb_env_tick(fs, n.adsr);
	F.in = osc;
	const float y0 = kb_onepole_tick(F.feedback, F.k * F.z[3]) * 0.9f * F.resonance;
	const float shaped = (y0 > KB_ROOT2_F) ? KB_ROOT2_F : (y0 < -0.5) ? -0.5f : y0;
	F.in -= shaped;
	F.z[0] += 2.f * F.b0 * (F.in - F.z[0] + F.z[1]);
	F.z[1] += F.b0 * (F.z[0] - 2.f * F.z[1] + F.z[2]);
	F.z[2] += F.b0 * (F.z[1] - 2.f * F.z[2] + F.z[3]);
	F.z[3] += F.b0 * (F.z[2] - 2.f * F.z[3]);
	const float x = F.g * F.z[3];
	const float hard = (x > 1.f) ? 1.f : (x < -1.f) ? -1.f : x;
	F.out = (float)((double)kb_tanhf(hard * F.drive) / B.clip_den);
	const float out = F.out * a;
	if (n.adsr.stage == KB_ENV_OFF) note_stage = KB_NOTE_OFF;
	return out;
}

// ---- effects, one frame each
// Gain.k:15-18
KB_D float kb_gain_frame(const KbFxHdr& h, float in) { return in * h.controls[0].value; }

// PingPong.k:42-59: the control half of the frame (smoothers, LFO, delay hand-over)
KB_D void kb_pingpong_control(const KbFs& fs, KbFxHdr& h, KbPingPong& p, float& gain, float& delay, float& dry) {
	KbControl* c = h.controls;
	const float rate = (c[3].value * c[3].value) * 100.f;                     // :44
	const float new_delay = kb_control_smooth(c[5]);                          // :45
	if (fabsf(p.delay - new_delay) > 0.001) {                                 // :46 (double compare)
		p.delay = new_delay;
		kb_control_set(c[1], new_delay);
		kb_bosc_set_fp(fs, p.lfo, rate, KB_PI_F);
	} else {
		p.delay = c[5].value;
		kb_bosc_set_f(fs, p.lfo, rate);
	}
	gain = c[0].value;                                                        // :55
	delay = kb_control_smooth(c[1]);                                          // :56
	const float vibrato = (c[2].value * c[2].value) * rate * KB_ROOT2_F;      // :57
	dry = c[4].value;                                                         // :58
	kb_control_set(c[1], c[1].value + kb_bosc_sine_tick(p.lfo) * vibrato * (float)0.00005);   // :59
}
// PingPong.k:42-71
KB_D void kb_pingpong_frame(const KbFs& fs, KbFxHdr& h, KbPingPong& p, float* rings, float inl, float inr, float& ol, float& orr) {
	float* ringl = rings + p.left.ring; float* ringr = rings + p.right.ring;
	float gain, delay, dry;
	kb_pingpong_control(fs, h, p, gain, delay, dry);
	kb_delay_set(p.left, delay * fs.f);                                       // :63
	kb_delay_set(p.right, 0.5f * delay * fs.f);                               // :64
	const float rr = kb_delay_tick(p.right, ringr);                           // :66 (Q13: two read ticks per line per frame)
	kb_delay_write(p.left, ringl, inl + rr * gain);
	const float lr = kb_delay_tick(p.left, ringl);
	const float outl = dry * inl + (1.f - dry) * lr;
	const float lr2 = kb_delay_tick(p.left, ringl);                           // :67
	kb_delay_write(p.right, ringr, inr + lr2 * gain);
	const float rr2 = kb_delay_tick(p.right, ringr);
	const float outr = dry * inr + (1.f - dry) * rr2;
	ol = kb_biquad_tick(p.dc[0], outl);                                       // :69
	orr = kb_biquad_tick(p.dc[1], outr);                                      // :70
}

#endif  // __CUDACC__
// (Reverb.k's frame uses no device-only math: also compiled by g++ for tests/host/reverb_scan_check.cpp)
// FilteredDelay::process  Reverb.k:130-132
KB_D float kb_rv_fdelay_tick(KbRvFDelay& d, float* rings) {
	float* ring = rings + d.delay.ring;
	kb_delay_write(d.delay, ring, d.in);
	const float x = kb_delay_tick(d.delay, ring);
	d.out = kb_biquad_tick(d.filter, x) * d.gain;
	return d.out;
}
// LateReflections::process: every FilteredDelay ticks twice per frame (Q12)   Reverb.k:152-169
KB_D float kb_rv_late_tick(KbRvLate& L, float* rings, float in) {
	L.in = in;
	float dl[4];
	for (int i = 0; i < 4; i++) dl[i] = kb_rv_fdelay_tick(L.d[i], rings);
	const float M[4][4] = { { 0, 1, 1,-1 }, {-1, 0,-1, 1 }, {-1, 1, 0,-1 }, { 1,-1, 1, 0 } };
	float fb[4];
	for (int r = 0; r < 4; r++)
		fb[r] = (M[r][0] * dl[0] + M[r][1] * dl[1] + M[r][2] * dl[2] + M[r][3] * dl[3]) + L.in;
	for (int i = 0; i < 4; i++) L.d[i].in = fb[i];
	float sum = kb_rv_fdelay_tick(L.d[0], rings);
	sum = sum + kb_rv_fdelay_tick(L.d[1], rings);
	sum = sum + kb_rv_fdelay_tick(L.d[2], rings);
	sum = sum + kb_rv_fdelay_tick(L.d[3], rings);
	L.out = sum;
	return sum;
}
// Reverb.k:266-279 with EarlyReflections::process (:86-92) and Reflections::process (:212-231)
KB_D void kb_reverb_frame(KbFxHdr& h, KbReverb& rv, float* rings, float inl, float inr, float& ol, float& orr) {
	const KbControl* c = h.controls;
	const float dry = c[0].value, wet = c[4].value;
	const float fl = kb_biquad_tick(rv.hpf[0], kb_biquad_tick(rv.lpf[0], inl));
	const float fr = kb_biquad_tick(rv.hpf[1], kb_biquad_tick(rv.lpf[1], inr));
	float* ringl = rings + rv.dl.ring; float* ringr = rings + rv.dr.ring;
	kb_delay_write(rv.dl, ringl, fl); kb_delay_write(rv.dr, ringr, fr);
	float r1l = 0, r1r = 0;
	for (int d = 0; d < rv.count; d++) {
		float tl, tr;
		kb_sdelay_tap_f(rv.dl, ringl, ringr, rv.times[d], tl, tr);
		r1l += tl * rv.gl[d];
		r1r += tr * rv.gr[d];
	}
	const float r2l = kb_rv_late_tick(rv.mid[0], rings, r1l);
	const float r2r = kb_rv_late_tick(rv.mid[1], rings, r1r);
	const float r3l = kb_rv_late_tick(rv.late[0], rings, r2l);
	const float r3r = kb_rv_late_tick(rv.late[1], rings, r2r);
	const float refl_l = (r1l * c[1].value + r2l * c[2].value) + r3l * c[3].value;
	const float refl_r = (r1r * c[1].value + r2r * c[2].value) + r3r * c[3].value;
	ol = inl * dry + refl_l * wet;                 // Reverb.k:272: `(in >> reflections) * wet` is signals<2>{wet, 0} (Q7)
	orr = inr * dry + refl_r * 0.f;
}
#ifdef __CUDACC__
// Delay/PingPong.k:24-34
KB_D void kb_dpingpong_frame(const KbFs& fs, KbFxHdr& h, KbDPingPong& p, float* rings, float inl, float inr, float& ol, float& orr) {
	const KbControl* c = h.controls;
	float* ringl = rings + p.l.ring; float* ringr = rings + p.r.ring;
	const float tl = c[0].value * fs.f, tr = c[1].value * fs.f;
	const float fl = kb_delay_tap_f(p.l, ringl, tl) * c[1].value;
	const float fr = kb_delay_tap_f(p.r, ringr, tr) * c[3].value;
	ol = inl + fr; orr = inr + fl;
	kb_delay_write(p.l, ringl, ol); kb_delay_write(p.r, ringr, orr);
}
// Delay/Reverb.k:59-78 (the double literals of :46-47 converted to float params)
KB_D float kb_dreverb_frame(const KbFs& fs, KbFxHdr& h, KbDReverb& p, float* rings, float in) {
	const float times[8] = { (float)2.078, (float)5.154, (float)5.947, (float)7.544, (float)8.878, (float)10.422, (float)13.938, (float)17.140 };
	const float gains[8] = { (float).609, (float).262, (float)-.360, (float)-.470, (float).290, (float)-.423, (float).100, (float).200 };
	const KbControl* c = h.controls;
	float* ff = rings + p.feedforward.ring; float* fb = rings + p.feedback.ring;
	kb_delay_write(p.feedforward, ff, in);
	float mix = in;
	for (int d = 0; d < 8; d++) mix += kb_delay_tap_f(p.feedforward, ff, times[d] * fs.f / 1000) * gains[d];
	const float late = kb_biquad_tick(p.filter, c[0].value * kb_delay_tap_f(p.feedback, fb, c[1].value * fs.f));
	p.out = mix + late;
	kb_delay_write(p.feedback, fb, p.out);
	return p.out;
}
#endif  // __CUDACC__
