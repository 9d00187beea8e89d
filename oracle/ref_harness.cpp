// TEST INFRASTRUCTURE — not part of the product.
//
// C-ABI harness around the UNMODIFIED reference implementation (nashaudio/klang,
// klang.h v0.7.8 and examples/*.k).  oracle/build_ref.py compiles this file
// against a build-time patched copy of /root/reference/klang.h (the one
// mechanical g++ patch described in oracle/build_ref.py) into
// oracle/_ref/libklang_ref.so.  Nothing here re-implements klang arithmetic: it
// only instantiates reference objects, drives them through their public API
// (set()/process()/start()/release()/Synth::process()/Effect::process()) and
// copies samples out, so it is the ground truth the restatement in
// oracle/klang_port.c and the CUDA path are checked against.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load the resulting library.
//
// Reference state is process-global (klang::fs, rand(), debug.buffer, the abs/sqr
// Function objects — klang.h:1593-1604, 3063-3069, 3196), so this library is
// single-threaded by contract; parallel CPU timing uses fork()ed processes.

#include <klang.h>

#include <cstdint>
#include <cstring>
#include <vector>

// ---------------------------------------------------------------------------
// The example programs, each inside its own namespace so that their global
// `using namespace klang::{optimised,basic}` directives and struct names do
// not collide.  klang.h is already included above (#pragma once).
// ---------------------------------------------------------------------------
namespace k_gain {
#include "Gain/Gain.k"
}
namespace k_filter {
#include "Subtractive/Filter.k"
}
namespace k_supersaw {
#include "SuperSaw.k"
}
namespace k_pingpong {
#include "PingPong.k"
}
namespace k_reverb {
#include "Reverb.k"
}
namespace k_tb303 {
#include "TB303_patched.k"   // one-token g++ disambiguation, see build_ref.py
}
namespace k_synthx {
#include "SynTHX_patched.k"  // one-token g++ disambiguation, see build_ref.py
}
namespace k_fm {
#include "FM.k"
}
namespace k_breakpoint {
#include "Subtractive/Breakpoint.k"
}
namespace k_ramp {
#include "Subtractive/Ramp.k"
}
namespace k_release {
#include "Subtractive/Release.k"
}
namespace k_pan {
#include "Gain/Pan.k"
}
namespace k_rm {
#include "Gain/RM.k"
}
namespace k_tremolo {
#include "Gain/Tremolo.k"
}
namespace k_clipping {
#include "Distortion/Clipping.k"
}
namespace k_add_saw {
#include "Additive/Saw.k"
}
namespace k_add_square {
#include "Additive/Square.k"
}
namespace k_am {
#include "Modulation/AM.k"
}
namespace k_mod_fm {
#include "Modulation/FM.k"
}
namespace k_mod_fm2 {
#include "Modulation/FM2.k"
}
namespace k_functions {
#include "Distortion/Functions.k"
}
namespace k_mute {
#include "Distortion/Mute.k"
}
namespace k_add_nyquist {
#include "Additive/Nyquist.k"
}
namespace k_iir {
#include "Filtering/IIR.k"
}
namespace k_wahwah {
#include "Filtering/WahWah.k"
}
namespace k_flanger {
#include "Modulation/Flanger.k"
}
namespace k_moddelay {
#include "Modulation/ModDelay.k"
}
namespace k_mod_chorus {
#include "Modulation/Chorus.k"
}
namespace k_echo {
#include "Delay/Echo.k"
}
namespace k_feedback {
#include "Delay/Feedback.k"
}
namespace k_delay_pingpong {
#include "Delay/PingPong.k"
}
namespace k_delay_reverb {
#include "Delay/Reverb.k"
}
// programs the product has NO hand-written graph for: it runs them from their own source (klang_b200/kcc.py); ids from 100
namespace k_objects {
#include "Filtering/Objects.k"
}
namespace k_bands {
#include "Filtering/Bands.k"
}
namespace k_eq {
#include "Filtering/EQ.k"
}
namespace k_patterns {
#include "Delay/Patterns.k"
}
namespace k_reverb2 {
#include "Delay/Reverb2.k"
}
namespace k_shaping {
#include "Distortion/Shaping.k"
}
namespace k_expression {
#include "Subtractive/Expression.k"
}
namespace k_resynthesis {
#include "Additive/Resynthesis.k"
}
namespace k_operators {
#include "Modulation/Operators.k"
}

// ---------------------------------------------------------------------------
// Canonical C2 graph: examples/Subtractive/Filter.k's note with
// `Square osc` replaced by `Saw osc`, ADSR times taken from four controls
// (defaults 0.01, 0.1, 0.7, 0.25).  Written in the reference DSL so the
// reference header evaluates it.
// ---------------------------------------------------------------------------
namespace k_subtractive {
using namespace klang::optimised;

struct Subtractive : Synth {
	struct SubNote : public Note {
		Saw osc;
		ADSR adsr;
		Envelope env;
		LPF filter;

		event on(Pitch pitch, Amplitude velocity) {
			param f = pitch -> Frequency;
			osc(f, 0);
			adsr(controls[0], controls[1], controls[2], controls[3]);
			env = { { 0, f * 2 }, { 0.25, f * 10 }, { 2, f * 5 } };
			filter.reset();
		}

		event off(Amplitude velocity) {
			adsr.release();
		}

		void process() {
			filter.set(env++, 10);
			osc >> filter >> out;

			out *= adsr++;
			if (adsr.finished())
				stop();
		}
	};

	Subtractive() {
		controls = {
			Dial("Attack", 0.0, 1.0, 0.01),
			Dial("Decay", 0.0, 1.0, 0.1),
			Dial("Sustain", 0.0, 1.0, 0.7),
			Dial("Release", 0.0, 1.0, 0.25),
		};
	}
};
}

using klang::Debug;

extern "C" {

// --------------------------------------------------------------------- globals
void ref_set_fs(float fs) { klang::fs = klang::SampleRate(fs); }
float ref_get_fs() { return klang::fs.f; }
void ref_srand(unsigned seed) { srand(seed); }
int ref_version() { return klang::version.major * 10000 + klang::version.minor * 100 + klang::version.build; }

float ref_pitch_to_frequency(float pitch) {
	klang::Pitch p(pitch);
	klang::param f = p -> Frequency;
	return f;
}

// ------------------------------------------------------------------ oscillators
enum { OSC_FAST_SAW = 0, OSC_FAST_TRIANGLE, OSC_FAST_SQUARE, OSC_FAST_PULSE, OSC_FAST_SINE,
       OSC_BASIC_SINE, OSC_BASIC_SAW, OSC_BASIC_TRIANGLE, OSC_BASIC_SQUARE, OSC_BASIC_PULSE,
       OSC_WT_SINE, OSC_WT_SAW, OSC_BASIC_NOISE, OSC_FAST_NOISE };

extern "C++" {
template<class OSC>
static void run_osc(int nargs, float f, float phase, float duty, int n, float* out) {
	OSC osc;
	if (nargs == 1) osc(klang::param(f));
	else if (nargs == 2) osc(klang::param(f), klang::param(phase));
	else if constexpr (std::is_base_of_v<klang::Generators::Fast::Osm, OSC> || std::is_same_v<OSC, klang::Generators::Basic::Pulse>)
		osc(klang::param(f), klang::param(phase), klang::param(duty));
	for (int s = 0; s < n; s++) {
		klang::signal y = osc;   // conversion ticks process() (klang.h:2214,2264)
		out[s] = y;
	}
}
} // extern "C++"

// nargs = number of set() arguments: 1 set(f), 2 set(f,phase), 3 set(f,phase,duty)
int ref_osc(int kind, int nargs, float f, float phase, float duty, int n, float* out) {
	using namespace klang::Generators;
	switch (kind) {
	case OSC_FAST_SAW:       run_osc<Fast::Saw>(nargs, f, phase, duty, n, out); break;
	case OSC_FAST_TRIANGLE:  run_osc<Fast::Triangle>(nargs, f, phase, duty, n, out); break;
	case OSC_FAST_SQUARE:    run_osc<Fast::Square>(nargs, f, phase, duty, n, out); break;
	case OSC_FAST_PULSE:     run_osc<Fast::Pulse>(nargs, f, phase, duty, n, out); break;
	case OSC_FAST_SINE:      run_osc<Fast::Sine>(nargs, f, phase, duty, n, out); break;
	case OSC_BASIC_SINE:     run_osc<Basic::Sine>(nargs, f, phase, duty, n, out); break;
	case OSC_BASIC_SAW:      run_osc<Basic::Saw>(nargs, f, phase, duty, n, out); break;
	case OSC_BASIC_TRIANGLE: run_osc<Basic::Triangle>(nargs, f, phase, duty, n, out); break;
	case OSC_BASIC_SQUARE:   run_osc<Basic::Square>(nargs, f, phase, duty, n, out); break;
	case OSC_BASIC_PULSE:    run_osc<Basic::Pulse>(nargs, f, phase, duty, n, out); break;
	case OSC_WT_SINE:        run_osc<Wavetables::Sine>(nargs, f, phase, duty, n, out); break;
	case OSC_WT_SAW:         run_osc<Wavetables::Saw>(nargs, f, phase, duty, n, out); break;
	// Noise has no set(): each tick is one libc rand() of the process-wide stream (klang.h:4947-4951, 5357-5366)
	case OSC_BASIC_NOISE:    { Basic::Noise o; for (int s = 0; s < n; s++) { klang::signal y = o; out[s] = y; } } break;
	case OSC_FAST_NOISE:     { Fast::Noise o; for (int s = 0; s < n; s++) { klang::signal y = o; out[s] = y; } } break;
	default: return -1;
	}
	return 0;
}

// klang::Sample (klang.h:3679-3720) attached to a caller's table: set(f) / set(f, phase) then n ticks.  (Phase::operator+= wraps on
// `position > size`, klang.h:1527-1534, so a run that reaches position == size reads samples[size]: callers stop before that.)
int ref_sample(const float* table, int size, int nargs, float f, float phase, int n, float* out) {
	klang::buffer buf(const_cast<float*>(table), size);
	klang::Sample smp;
	smp = buf;
	if (nargs == 1) smp(klang::param(f)); else smp(klang::param(f), klang::param(phase));
	for (int s = 0; s < n; s++) { klang::signal y = smp; out[s] = y; }
	return 0;
}
// File::WAV (klang.h:5951-6085): load(path) walks the RIFF chunks, operator>> decodes data->size / BlockAlign samples (8-bit unsigned, 16 / 32-bit
// signed PCM, 32-bit float) into a variable::buffer.  info = { NumChannels, SampleRate, BitsPerSample }.  Returns the number of samples decoded.
int ref_wav_decode(const char* path, float* out, int max, int* info) {
	klang::File::WAV wav;
	if (!wav.load(path)) return -1;
	klang::variable::buffer buf;
	if (!(wav >> buf)) return -2;
	if (info) { info[0] = wav.format->NumChannels; info[1] = (int)wav.format->SampleRate; info[2] = wav.format->BitsPerSample; }
	for (int i = 0; i < buf.size && i < max; i++) out[i] = buf.data()[i];
	return buf.size;
}

// Wavetable contents (2048 floats) as the reference builds them (klang.h:3645-3650, 5372-5379).
int ref_wavetable(int kind, float* table) {
	using namespace klang::Generators;
	if (kind == OSC_WT_SINE) { Wavetables::Sine w; for (int i = 0; i < 2048; i++) table[i] = w[i]; return 0; }
	if (kind == OSC_WT_SAW)  { Wavetables::Saw w;  for (int i = 0; i < 2048; i++) table[i] = w[i]; return 0; }
	return -1;
}

// ---------------------------------------------------------------------- filters
enum { FLT_BIQUAD_LPF = 0, FLT_BIQUAD_HPF, FLT_ONEPOLE_LPF, FLT_ONEPOLE_HPF,
       FLT_BIQUAD_BPF, FLT_BIQUAD_BRF, FLT_BIQUAD_APF, FLT_BUTTERWORTH_LPF1, FLT_BUTTERWORTH_LPF2,
       FLT_DCF, FLT_IIR1, FLT_IIR2, FLT_MODAL, FLT_FOLLOWER_PEAK, FLT_FOLLOWER_RMS, FLT_WINDOW_MEAN, FLT_WINDOW_RMS };

extern "C++" {
template<class F>
static void run_biquad(int nset, const float* f, const float* Q, int n, const float* in, float* out, float* coeffs) {
	F flt;
	for (int s = 0; s < n; s++) {
		if (s < nset) {
			if (Q) flt.set(klang::param(f[s]), klang::param(Q[s]));
			else flt.set(klang::param(f[s]));
		}
		klang::signal x = in[s];
		klang::signal y;
		x >> flt >> y;
		out[s] = y;
	}
	if (coeffs) { coeffs[0] = flt.b0; coeffs[1] = flt.b1; coeffs[2] = flt.b2; coeffs[3] = flt.a1; coeffs[4] = flt.a2; }
}

template<class F>
static void run_onepole(int nset, const float* f, int n, const float* in, float* out, float* coeffs) {
	F flt;
	for (int s = 0; s < n; s++) {
		if (s < nset) flt.set(klang::param(f[s]));
		klang::signal x = in[s];
		klang::signal y;
		x >> flt >> y;
		out[s] = y;
	}
	if (coeffs) { coeffs[0] = flt.b0; coeffs[1] = flt.b1; coeffs[2] = 0; coeffs[3] = flt.a1; coeffs[4] = 0; }
}
// Filters::DCF (f = r), IIR<1> (f = coefficient), IIR<2> (f = a1, Q = a2)   klang.h:5387-5464
template<class F, class SET>
static void run_simple(int nset, int n, const float* in, float* out, F& flt, SET set) {
	for (int s = 0; s < n; s++) {
		if (s < nset) set(s);
		klang::signal x = in[s];
		klang::signal y;
		x >> flt >> y;
		out[s] = y;
	}
}
} // extern "C++"

// nset: how many leading samples call set(f[s](,Q[s])) before processing (1 = static, n = per-sample).
// Q == NULL → set(f) (default Q = 1/sqrt2 for biquads).
int ref_filter(int kind, int nset, const float* f, const float* Q, int n, const float* in, float* out, float* coeffs) {
	using namespace klang::Filters;
	switch (kind) {
	case FLT_BIQUAD_LPF: run_biquad<Biquad::LPF>(nset, f, Q, n, in, out, coeffs); break;
	case FLT_BIQUAD_HPF: run_biquad<Biquad::HPF>(nset, f, Q, n, in, out, coeffs); break;
	case FLT_BIQUAD_BPF: run_biquad<Biquad::BPF>(nset, f, Q, n, in, out, coeffs); break;
	case FLT_BIQUAD_BRF: run_biquad<Biquad::BRF>(nset, f, Q, n, in, out, coeffs); break;
	case FLT_BIQUAD_APF: run_biquad<Biquad::APF>(nset, f, Q, n, in, out, coeffs); break;
	case FLT_BUTTERWORTH_LPF2: run_biquad<Butterworth::LPF<2>>(nset, f, Q, n, in, out, coeffs); break;
	case FLT_ONEPOLE_LPF: run_onepole<OnePole::LPF>(nset, f, n, in, out, coeffs); break;
	case FLT_ONEPOLE_HPF: run_onepole<OnePole::HPF>(nset, f, n, in, out, coeffs); break;
	case FLT_BUTTERWORTH_LPF1: run_onepole<Butterworth::LPF<1>>(nset, f, n, in, out, coeffs); break;
	case FLT_DCF: { DCF flt; run_simple(nset, n, in, out, flt, [&](int s) { flt.set(f[s]); }); if (coeffs) { coeffs[0] = flt.r; coeffs[1] = flt.z; coeffs[2] = coeffs[3] = coeffs[4] = 0; } break; }
	case FLT_IIR1: { IIR<1> flt; run_simple(nset, n, in, out, flt, [&](int s) { flt.set(klang::param(f[s])); }); if (coeffs) { coeffs[0] = flt.a; coeffs[1] = flt.b; coeffs[2] = coeffs[3] = coeffs[4] = 0; } break; }
	case FLT_IIR2: { if (!Q) return -1; IIR<2> flt; run_simple(nset, n, in, out, flt, [&](int s) { flt.set(f[s], Q[s]); }); if (coeffs) { coeffs[0] = flt.a[0]; coeffs[1] = flt.a[1]; coeffs[2] = flt.y[0]; coeffs[3] = flt.y[1]; coeffs[4] = 0; } break; }
	// Modifiers::Modal (f, Q = decay seconds)  klang.h:5817-5857;  Envelope::Follower peak / rms (f = attack, Q = release)  5862-5896
	case FLT_MODAL: { if (!Q) return -1; klang::Modifiers::Modal flt; run_simple(nset, n, in, out, flt, [&](int s) { flt.set(klang::param(f[s]), klang::param(Q[s])); });
		if (coeffs) { coeffs[0] = flt.a1; coeffs[1] = flt.a2; coeffs[2] = flt.gain; coeffs[3] = flt.y1; coeffs[4] = flt.y2; } break; }
	case FLT_FOLLOWER_PEAK: case FLT_FOLLOWER_RMS: { if (!Q) return -1; klang::Envelope::Follower flt; flt = (kind == FLT_FOLLOWER_RMS) ? klang::RMS : klang::Peak;
		run_simple(nset, n, in, out, flt, [&](int s) { flt.set(klang::param(f[s]), klang::param(Q[s])); });
		if (coeffs) { coeffs[0] = flt.ar.A; coeffs[1] = flt.ar.R; coeffs[2] = flt.ar.out; coeffs[3] = coeffs[4] = 0; } break; }
	// Envelope::Follower::Window<64> (klang.h:5904-5948): 64-sample moving sum kept in a double, then the attack / release smoother
	case FLT_WINDOW_MEAN: case FLT_WINDOW_RMS: { if (!Q) return -1; klang::Envelope::Follower::Window<64> flt; flt = (kind == FLT_WINDOW_RMS) ? klang::RMS : klang::Mean;
		run_simple(nset, n, in, out, flt, [&](int s) { flt.set(klang::param(f[s]), klang::param(Q[s])); });
		if (coeffs) { coeffs[0] = flt.ar.A; coeffs[1] = flt.ar.R; coeffs[2] = flt.ar.out; coeffs[3] = (float)flt.sum; coeffs[4] = 0; } break; }
	default: return -1;
	}
	return 0;
}

// -------------------------------------------------------------------- envelopes
// points: npts (x,y) pairs.  loop_start/loop_end = -1 for none.  release_at: sample index at which
// release(release_time, release_level) is called (-1 never).  stage_out[s] = stage after sample s.
int ref_envelope(int npts, const float* xy, int loop_start, int loop_end, int n, int release_at,
                 float release_time, float release_level, float* out, int* stage_out) {
	std::vector<klang::Envelope::Point> pts;
	for (int p = 0; p < npts; p++) pts.push_back({ xy[2 * p], xy[2 * p + 1] });
	klang::Envelope env;
	env.set(pts);
	if (loop_start >= 0) env.setLoop(loop_start, loop_end);
	for (int s = 0; s < n; s++) {
		if (s == release_at) env.release(release_time, release_level);
		out[s] = env++;
		if (stage_out) stage_out[s] = (int)env.getStage();
	}
	return 0;
}

int ref_envelope_at(int npts, const float* xy, int n, const float* t, float* out) {
	std::vector<klang::Envelope::Point> pts;
	for (int p = 0; p < npts; p++) pts.push_back({ xy[2 * p], xy[2 * p + 1] });
	klang::Envelope env;
	env.set(pts);
	for (int s = 0; s < n; s++) out[s] = env.at(t[s]);
	return 0;
}

int ref_adsr(float A, float D, float S, float R, int n, int release_at, float* out, int* stage_out) {
	klang::ADSR adsr;
	adsr(klang::param(A), klang::param(D), klang::param(S), klang::param(R));
	for (int s = 0; s < n; s++) {
		if (s == release_at) adsr.release();
		out[s] = adsr++;
		if (stage_out) stage_out[s] = (int)adsr.getStage();
	}
	return 0;
}

// ----------------------------------------------------------------------- delays
// Mono Delay<1000>: per sample s: write in[s]; then out_i[s] = tap(int di[s]); out_f[s] = tap(float df[s]);
// if set_at[s] >= 0: set(set_at[s]); out_p[s] = process() result (read head), only valid after first set.
int ref_delay1000(int n, const float* in, const int* di, const float* df, const float* set_at,
                  float* out_i, float* out_f, float* out_p) {
	klang::Delay<1000>* d = new klang::Delay<1000>();
	bool have_set = false;
	for (int s = 0; s < n; s++) {
		klang::signal x = in[s];
		x >> *d;
		out_i[s] = d->tap(di[s]);
		out_f[s] = d->tap(df[s]);
		if (set_at[s] >= 0.f) { d->set(klang::param(set_at[s])); have_set = true; }
		if (have_set) { d->process(); out_p[s] = d->out; } else out_p[s] = 0.f;
	}
	delete d;
	return 0;
}

// Delay<1000>::lagrange(df[s]) after writing in[s] (third-order interpolation, klang.h:3429-3458)
int ref_delay1000_lagrange(int n, const float* in, const float* df, float* out) {
	klang::Delay<1000>* d = new klang::Delay<1000>();
	for (int s = 0; s < n; s++) {
		klang::signal x = in[s];
		x >> *d;
		out[s] = d->lagrange(df[s]);
	}
	delete d;
	return 0;
}

// Stereo::Delay<1000>: per sample write (inl,inr); out = tap(float df[s]) (klang.h:4668-4681).
int ref_stereo_delay1000(int n, const float* inl, const float* inr, const float* df, float* outl, float* outr) {
	klang::Stereo::Delay<1000>* d = new klang::Stereo::Delay<1000>();
	for (int s = 0; s < n; s++) {
		klang::Stereo::signal x(inl[s], inr[s]);
		x >> *d;
		klang::Stereo::signal y = (*d)(df[s]);
		outl[s] = y.l; outr[s] = y.r;
	}
	delete d;
	return 0;
}

// Control::smooth / Control::set (klang.h:1715-1728)
int ref_control_smooth(float lo, float hi, float initial, int n, const float* values, float* out) {
	klang::Control c = klang::Dial("c", lo, hi, initial);
	for (int s = 0; s < n; s++) {
		c.set(values[s]);
		out[s] = c.smooth();
	}
	return 0;
}

// ---------------------------------------------------------------------- effects
enum { FX_GAIN = 0, FX_PINGPONG = 1, FX_REVERB = 2, FX_DELAY_PINGPONG = 3, FX_DELAY_REVERB = 4, FX_PAN = 5, FX_RM = 6, FX_TREMOLO = 7, FX_CLIPPING = 8, FX_ECHO = 9, FX_FEEDBACK = 10, FX_FUNCTIONS = 11, FX_MUTE = 12, FX_IIR = 13, FX_WAHWAH = 14, FX_FLANGER = 15, FX_MODDELAY = 16, FX_MOD_CHORUS = 17 };

struct RefFx {
	int graph;
	klang::Effect* mono = nullptr;
	klang::Stereo::Effect* stereo = nullptr;
	klang::Controls* controls = nullptr;
};

void* ref_fx_create(int graph) {
	RefFx* fx = new RefFx;
	fx->graph = graph;
	switch (graph) {
	case FX_GAIN:     { auto* e = new k_gain::Gain();         fx->mono = e;   fx->controls = &e->controls; } break;
	case FX_PINGPONG: { auto* e = new k_pingpong::PingPong(); fx->stereo = e; fx->controls = &e->controls; } break;
	case FX_REVERB:   { auto* e = new k_reverb::Reverb();     fx->stereo = e; fx->controls = &e->controls; } break;
	case FX_DELAY_PINGPONG: { auto* e = new k_delay_pingpong::PingPong(); fx->stereo = e; fx->controls = &e->controls; } break;
	case FX_DELAY_REVERB:   { auto* e = new k_delay_reverb::Reverb();     fx->mono = e;   fx->controls = &e->controls; } break;
	case FX_PAN:      { auto* e = new k_pan::Pan();           fx->stereo = e; fx->controls = &e->controls; } break;
	case FX_RM:       { auto* e = new k_rm::RM();             fx->mono = e;   fx->controls = &e->controls; } break;
	case FX_TREMOLO:  { auto* e = new k_tremolo::Tremolo();   fx->mono = e;   fx->controls = &e->controls; } break;
	case FX_CLIPPING: { auto* e = new k_clipping::Clipping(); fx->mono = e;   fx->controls = &e->controls; } break;
	case FX_ECHO:     { auto* e = new k_echo::Echo();         fx->mono = e;   fx->controls = &e->controls; } break;
	case FX_FEEDBACK: { auto* e = new k_feedback::Feedback(); fx->mono = e;   fx->controls = &e->controls; } break;
	case FX_FUNCTIONS: { auto* e = new k_functions::Functions(); fx->mono = e; fx->controls = &e->controls; } break;
	case FX_MUTE:     { auto* e = new k_mute::Mute();         fx->mono = e;   fx->controls = &e->controls; } break;
	case FX_IIR:      { auto* e = new k_iir::IIR();           fx->mono = e;   fx->controls = &e->controls; } break;
	case FX_WAHWAH:   { auto* e = new k_wahwah::WahWah();     fx->mono = e;   fx->controls = &e->controls; } break;
	case FX_FLANGER:  { auto* e = new k_flanger::Flanger();   fx->mono = e;   fx->controls = &e->controls; } break;
	case FX_MODDELAY: { auto* e = new k_moddelay::ModDelay(); fx->mono = e;   fx->controls = &e->controls; } break;
	case FX_MOD_CHORUS: { auto* e = new k_mod_chorus::Chorus(); fx->mono = e; fx->controls = &e->controls; } break;
	case 100: { auto* e = new k_objects::Objects(); fx->mono = e; fx->controls = &e->controls; } break;     // Filtering/Objects.k (Noise >> LPF)
	case 101: { auto* e = new k_bands::Bands();     fx->mono = e; fx->controls = &e->controls; } break;     // Filtering/Bands.k (two BPF, grouped controls)
	case 102: { auto* e = new k_eq::EQ();           fx->mono = e; fx->controls = &e->controls; } break;     // Filtering/EQ.k (LPF / HPF set in prepare())
	case 103: { auto* e = new k_patterns::Patterns(); fx->mono = e; fx->controls = &e->controls; } break;   // Delay/Patterns.k (a Menu control selects the tap pattern)
	case 104: { auto* e = new k_reverb2::Reverb2();  fx->stereo = e; fx->controls = &e->controls; } break;  // Delay/Reverb2.k (four Delay<192000>, two LPF, in[c] / out.l)
	case 105: { auto* e = new k_shaping::Shaping();  fx->mono = e;   fx->controls = &e->controls; } break;  // Distortion/Shaping.k (Function<float, float> over softclip: tanh(c x) / tanh(c))
	default: delete fx; return nullptr;
	}
	return fx;
}

void ref_fx_destroy(void* h) {
	RefFx* fx = (RefFx*)h;
	if (!fx) return;
	delete fx->mono;
	delete fx->stereo;
	delete fx;
}

int ref_fx_channels(void* h) { return ((RefFx*)h)->stereo ? 2 : 1; }
int ref_fx_num_controls(void* h) { return (int)((RefFx*)h)->controls->size(); }
void ref_fx_set_control(void* h, int idx, float v) { (*((RefFx*)h)->controls)[idx].set(v); }
float ref_fx_get_control(void* h, int idx) { return (*((RefFx*)h)->controls)[idx].value; }

// In-place block processing through Effect::process(buffer) (klang.h:4208-4216, 4708-4716).
int ref_fx_process(void* h, float* l, float* r, int n) {
	RefFx* fx = (RefFx*)h;
	if (n > 16384) return -1; // debug.buffer capacity (klang.h:3141)
	Debug::Session session(nullptr, n, Debug::Buffer::Effect);
	if (fx->mono) {
		klang::buffer b(l, n);
		fx->mono->process(b);
	} else {
		klang::buffer bl(l, n), br(r, n);
		klang::Stereo::buffer b(bl, br);
		fx->stereo->process(b);
	}
	return 0;
}

// The debug tap of the last block (`x >> debug`: Debug::input adds into Debug::buffer, klang.h:3132-3287; the block drivers step it,
// klang.h:4214 / 4300 / 4714; the host's Debug::Session clears it before the block, klang.h:3223-3245).  Returns 1 and copies n samples if
// the block wrote to it (Buffer::get, klang.h:3164-3172), else 0.
int ref_fx_debug(void* h, float* dst, int n) {
	(void)h;
	const float* p = Debug::buffer.get();
	if (!p) return 0;
	memcpy(dst, p, (size_t)n * sizeof(float));
	return 1;
}

// Factory presets (Plugin::presets, klang.h:1940-1981, 4195-4200) and the host's preset load: every value goes through Control::set like a
// host parameter (klang.h:1725-1728, 4444-4447), then Controller::onPreset (klang.h:4190, 4415-4420).
static int preset_get(klang::Plugin* pl, int p, char* name, int name_max, float* values, int max) {
	if (p < 0 || p >= (int)pl->presets.count) return -1;
	const klang::Preset& pr = pl->presets[p];
	if (name && name_max > 0) { strncpy(name, pr.name.c_str(), name_max - 1); name[name_max - 1] = 0; }
	const int nv = (int)pr.values.count;
	for (int c = 0; c < nv && c < max; c++) values[c] = pr.values[c];
	return nv;
}
static int preset_load(klang::Plugin* pl, int p) {
	if (p < 0 || p >= (int)pl->presets.count) return -1;
	const klang::Preset& pr = pl->presets[p];
	for (int c = 0; c < (int)pr.values.count && c < (int)pl->controls.size(); c++) pl->controls[c].set(pr.values[c]);
	pl->onPreset(p);
	return 0;
}
static klang::Plugin* fx_plugin(void* h) { RefFx* fx = (RefFx*)h; return fx->mono ? static_cast<klang::Plugin*>(fx->mono) : static_cast<klang::Plugin*>(fx->stereo); }
int ref_fx_num_presets(void* h) { return (int)fx_plugin(h)->presets.count; }
int ref_fx_preset(void* h, int p, char* name, int name_max, float* values, int max) { return preset_get(fx_plugin(h), p, name, name_max, values, max); }
int ref_fx_load_preset(void* h, int p) { return preset_load(fx_plugin(h), p); }

// ----------------------------------------------------------------------- synths
enum { SY_SUBTRACTIVE = 0, SY_SUPERSAW = 1, SY_TB303 = 2, SY_SYNTHX = 3, SY_FILTER_K = 4, SY_FM = 5, SY_BREAKPOINT = 6, SY_RAMP = 7, SY_RELEASE = 8, SY_ADDITIVE_SAW = 9, SY_ADDITIVE_SQUARE = 10, SY_AM = 11, SY_MOD_FM = 12, SY_MOD_FM2 = 13, SY_ADDITIVE_NYQUIST = 14 };

struct RefSynth {
	int graph;
	int nvoices;
	klang::Synth* mono = nullptr;
	klang::Stereo::Synth* stereo = nullptr;
	klang::Controls* controls = nullptr;
	std::vector<float> scratch;
};

extern "C++" {
template<class SYNTH, class NOTE>
static SYNTH* make_synth(int nvoices) {
	SYNTH* s = new SYNTH();
	int have = (int)s->notes.count;
	if (nvoices > have) s->notes.template add<NOTE>(nvoices - have);
	return s;
}
} // extern "C++"

void* ref_synth_create(int graph, int nvoices) {
	if (nvoices < 1 || nvoices > 128) return nullptr;
	RefSynth* s = new RefSynth;
	s->graph = graph;
	switch (graph) {
	case SY_SUBTRACTIVE: { auto* p = make_synth<k_subtractive::Subtractive, k_subtractive::Subtractive::SubNote>(nvoices); s->mono = p; s->controls = &p->controls; } break;
	case SY_SUPERSAW:    { auto* p = make_synth<k_supersaw::SuperSaw, k_supersaw::SuperSaw::MyNote>(nvoices); s->mono = p; s->controls = &p->controls; } break;
	case SY_TB303:       { auto* p = make_synth<k_tb303::TB303, k_tb303::TB303::MyNote>(nvoices); s->mono = p; s->controls = &p->controls; } break;
	case SY_FILTER_K:    { auto* p = make_synth<k_filter::Filter, k_filter::Filter::FilterNote>(nvoices); s->mono = p; s->controls = &p->controls; } break;
	case SY_SYNTHX:      { auto* p = make_synth<k_synthx::SynTHX, k_synthx::SynTHX::MyNote>(nvoices); s->stereo = p; s->controls = &p->controls; } break;
	case SY_FM:          { auto* p = make_synth<k_fm::FM, k_fm::FM::MyNote>(nvoices); s->mono = p; s->controls = &p->controls; } break;
	case SY_BREAKPOINT:  { auto* p = make_synth<k_breakpoint::Breakpoint, k_breakpoint::Breakpoint::BreakpointNote>(nvoices); s->mono = p; s->controls = &p->controls; } break;
	case SY_RAMP:        { auto* p = make_synth<k_ramp::Ramp, k_ramp::Ramp::RampNote>(nvoices); s->mono = p; s->controls = &p->controls; } break;
	case SY_RELEASE:     { auto* p = make_synth<k_release::Release, k_release::Release::ReleaseNote>(nvoices); s->mono = p; s->controls = &p->controls; } break;
	case SY_ADDITIVE_SAW:    { auto* p = make_synth<k_add_saw::Saw, k_add_saw::Saw::SawNote>(nvoices); s->mono = p; s->controls = &p->controls; } break;
	case SY_ADDITIVE_SQUARE: { auto* p = make_synth<k_add_square::Square, k_add_square::Square::SquareNote>(nvoices); s->mono = p; s->controls = &p->controls; } break;
	case SY_ADDITIVE_NYQUIST: { auto* p = make_synth<k_add_nyquist::Nyquist, k_add_nyquist::Nyquist::NyquistNote>(nvoices); s->mono = p; s->controls = &p->controls; } break;
	case SY_AM:      { auto* p = make_synth<k_am::AM, k_am::AM::AMNote>(nvoices); s->mono = p; s->controls = &p->controls; } break;
	case SY_MOD_FM:  { auto* p = make_synth<k_mod_fm::FM, k_mod_fm::FM::FMNote>(nvoices); s->mono = p; s->controls = &p->controls; } break;
	case SY_MOD_FM2: { auto* p = make_synth<k_mod_fm2::FM2, k_mod_fm2::FM2::FM2Note>(nvoices); s->mono = p; s->controls = &p->controls; } break;
	// programs the product has NO hand-written graph for (run from their own source, klang_b200/kcc.py): ids from 100
	case 100: { auto* p = make_synth<k_expression::Expression, k_expression::Expression::ExpressionNote>(nvoices); s->mono = p; s->controls = &p->controls; } break;   // Subtractive/Expression.k
	case 101: { auto* p = make_synth<k_resynthesis::Resynthesis, k_resynthesis::Resynthesis::ResynthesisNote>(nvoices); s->mono = p; s->controls = &p->controls; } break;   // Additive/Resynthesis.k
	case 102: { auto* p = make_synth<k_operators::Operators, k_operators::Operators::OperatorsNote>(nvoices); s->mono = p; s->controls = &p->controls; } break;   // Modulation/Operators.k
	default: delete s; return nullptr;
	}
	s->nvoices = s->mono ? (int)s->mono->notes.count : (int)s->stereo->notes.count;
	return s;
}

void ref_synth_destroy(void* h) {
	RefSynth* s = (RefSynth*)h;
	if (!s) return;
	delete s->mono;
	delete s->stereo;
	delete s;
}

static klang::Plugin* synth_plugin(void* h) { RefSynth* s = (RefSynth*)h; return s->mono ? static_cast<klang::Plugin*>(s->mono) : static_cast<klang::Plugin*>(s->stereo); }
int ref_synth_num_presets(void* h) { return (int)synth_plugin(h)->presets.count; }
int ref_synth_preset(void* h, int p, char* name, int name_max, float* values, int max) { return preset_get(synth_plugin(h), p, name, name_max, values, max); }
int ref_synth_load_preset(void* h, int p) { return preset_load(synth_plugin(h), p); }
// Synth::onControl (klang.h:4399-4404 / 4789-4794): the synth's control() hook, then every note that is not Off.  Returns how many notes that is.
int ref_synth_on_control(void* h, int idx, float value) {
	RefSynth* s = (RefSynth*)h;
	int notified = 0;
	if (s->mono) { s->mono->onControl(idx, value); for (unsigned i = 0; i < s->mono->notes.count; i++) notified += s->mono->notes[i]->stage != klang::Synth::Note::Off; }
	else { s->stereo->onControl(idx, value); for (unsigned i = 0; i < s->stereo->notes.count; i++) notified += s->stereo->notes[i]->stage != klang::Stereo::Synth::Note::Off; }
	return notified;
}

int ref_synth_channels(void* h) { return ((RefSynth*)h)->stereo ? 2 : 1; }
int ref_synth_num_voices(void* h) { return ((RefSynth*)h)->nvoices; }
int ref_synth_num_controls(void* h) { return (int)((RefSynth*)h)->controls->size(); }
void ref_synth_set_control(void* h, int idx, float v) { (*((RefSynth*)h)->controls)[idx].set(v); }
float ref_synth_get_control(void* h, int idx) { return (*((RefSynth*)h)->controls)[idx].value; }

// Synth::noteOn (klang.h:4423-4427); returns the voice index Notes::assign() picked.
int ref_synth_note_on(void* h, int pitch, float velocity) {
	RefSynth* s = (RefSynth*)h;
	if (s->mono) {
		s->mono->noteOn(pitch, velocity);
		for (unsigned i = 0; i < s->mono->notes.count; i++)
			if (s->mono->notes.noteStart[i] == s->mono->notes.noteOns - 1) return (int)i;
	} else {
		s->stereo->noteOn(pitch, velocity);
		for (unsigned i = 0; i < s->stereo->notes.count; i++)
			if (s->stereo->notes.noteStart[i] == s->stereo->notes.noteOns - 1) return (int)i;
	}
	return -1;
}

// Synth::noteOff (klang.h:4430-4434)
void ref_synth_note_off(void* h, int pitch, float velocity) {
	RefSynth* s = (RefSynth*)h;
	if (s->mono) s->mono->noteOff(pitch, velocity); else s->stereo->noteOff(pitch, velocity);
}

// Direct voice control: NoteBase::start / release (klang.h:4257-4275)
void ref_synth_voice_start(void* h, int voice, float pitch, float velocity) {
	RefSynth* s = (RefSynth*)h;
	if (s->mono) s->mono->notes[voice]->start(pitch, velocity); else s->stereo->notes[voice]->start(pitch, velocity);
}
void ref_synth_voice_release(void* h, int voice, float velocity) {
	RefSynth* s = (RefSynth*)h;
	if (s->mono) s->mono->notes[voice]->release(velocity); else s->stereo->notes[voice]->release(velocity);
}
// 0 Onset, 1 Sustain, 2 Release, 3 Off (klang.h:4284)
int ref_synth_voice_stage(void* h, int voice) {
	RefSynth* s = (RefSynth*)h;
	return s->mono ? (int)s->mono->notes[voice]->stage : (int)s->stereo->notes[voice]->stage;
}

// The reference block driver itself: Synth::process(float*,int,float*) (klang.h:4440-4466) or
// Stereo::Synth::process(float**,int,float*) (klang.h:4830-4858).  Buffers are cleared first
// (stereo notes accumulate, klang.h:4731).
int ref_synth_process(void* h, float* l, float* r, int n) {
	RefSynth* s = (RefSynth*)h;
	if (n > 16384) return -1;
	Debug::Session session(nullptr, n, Debug::Buffer::Synth);
	memset(l, 0, sizeof(float) * n);
	if (s->mono) {
		s->mono->klang::Synth::process(l, n, nullptr);
	} else {
		memset(r, 0, sizeof(float) * n);
		float* bufs[2] = { l, r };
		s->stereo->klang::Stereo::Synth::process(bufs, n, nullptr);
	}
	return 0;
}

// Per-voice rendering: every voice whose stage != Off is rendered ALONE through its public
// Note::process(buffer) into a zeroed buffer (out[v][c][n], c < channels); finished voices are
// stop()ped exactly like the Synth driver does (klang.h:4452-4455).  active[v] = 1 if rendered.
// No synth-level post-fx is applied.
int ref_synth_process_voices(void* h, float* out, int n, int* active) {
	RefSynth* s = (RefSynth*)h;
	if (n > 16384) return -1;
	const int C = s->stereo ? 2 : 1;
	for (int v = 0; v < s->nvoices; v++) {
		float* o = out + (size_t)v * C * n;
		memset(o, 0, sizeof(float) * C * n);
		if (s->mono) {
			klang::Note* note = s->mono->notes[v];
			active[v] = note->stage != klang::Note::Off;
			if (active[v]) {
				Debug::Session session(nullptr, n, Debug::Buffer::Synth);
				klang::buffer b(o, n);
				if (!note->process(b)) note->stop();
			}
		} else {
			klang::Stereo::Note* note = s->stereo->notes[v];
			active[v] = note->stage != klang::Stereo::Note::Off;
			if (active[v]) {
				Debug::Session session(nullptr, n, Debug::Buffer::Synth);
				klang::buffer bl(o, n), br(o + n, n);
				klang::Stereo::buffer b(bl, br);
				if (!note->process(b)) note->stop();
			}
		}
	}
	return 0;
}

} // extern "C"
