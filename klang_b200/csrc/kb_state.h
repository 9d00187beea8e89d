// klang-b200 — plain-old-data state of every primitive on the hot path.
//
// One struct per reference primitive, laid out so that a voice / effect instance is a flat POD blob:
// the host runs the event-rate code (Note::on/off, Effect::prepare, Control::set) on a mirror of the
// blob, the device runs the per-sample code (process()) on the copy in HBM.  Reference classes are
// cited per struct (klang.h = nashaudio/klang v0.7.8).
#pragma once
#include <stdint.h>

#define KB_MAX_CONTROLS 16
#define KB_MAX_VOICES 128      // Notes = Array<NOTE*,128>                 klang.h:4311
#define KB_ENV_MAXPTS 16

// klang::SampleRate                                                        klang.h:1593-1604
struct KbFs { float f; int i; float inv, w, nyquist; };

// klang::Control (value range + one-pole smoother)                         klang.h:1655-1755
struct KbControl { float min, max, value, smoothed; };

// Generators::Fast::OSM + Osm (uint32 phase, 6-state band-limited osc)     klang.h:5175-5354
struct KbOsm {
	int waveform;                       // 0 = saw()/triangle, 1 = pulse()/square
	int increment; uint32_t offset, duty; int state;
	float delta, f, omf, rcpf, rcpf2, col, c1, c2;
	float frequency;
};

// Generators::Fast::Sine                                                    klang.h:5135-5172
struct KbFastSine { float frequency; int increment; uint32_t position, offset; };

// Generic::Oscillator with float phase (Generators::Basic::*)               klang.h:2849-2880, 4899-4944
struct KbBasicOsc { float increment, position, frequency, offset, duty; };

// Filters::Biquad::Filter (TDF-II)                                          klang.h:5550-5612
struct KbBiquad { int type; float f, Q, a1, a2, b0, b1, b2, a, cos0, sin0, z0, z1; };

// Filters::OnePole::Filter / Butterworth::LPF<1>                            klang.h:5470-5543, 5786-5799
struct KbOnePole { int type; float f, a1, b0, b1, z, out; };

// Envelope + Envelope::Linear ramp (+ the ADSR parameters)                  klang.h:3723-4137
struct KbEnv {
	float px[KB_ENV_MAXPTS], py[KB_ENV_MAXPTS];
	int npoints, loop_start, loop_end, point;
	float time, timeInc; int stage;
	float out;
	float r_out, r_target, r_rate; int r_active;
	float A, D, S, R;
};

// Delay<SIZE>: the ring itself lives in the bank's HBM ring arena            klang.h:3381-3512
struct KbDelay { int SIZE; int position; int last_position; float last_fraction; float time; float out; long long ring; /* float offset into the arena */ };

// NoteBase                                                                   klang.h:4220-4290
enum { KB_NOTE_ONSET = 0, KB_NOTE_SUSTAIN = 1, KB_NOTE_RELEASE = 2, KB_NOTE_OFF = 3 };
enum { KB_ENV_SUSTAIN = 0, KB_ENV_RELEASE = 1, KB_ENV_OFF = 2 };
struct KbVoiceHdr { int stage; float pitch, velocity; int active; /* stage != Off when the block started */ };

// ------------------------------------------------------------------ synth voices (graphs)
// examples/Subtractive/Filter.k:7-36 and the canonical C2 graph
struct KbSubVoice { KbOsm osc; KbEnv adsr; KbEnv env; KbBiquad filter; };
// examples/SuperSaw.k:7-34
struct KbSsawVoice { KbOsm osc[7]; KbEnv adsr; };
// examples/TB303.k:8-114
struct KbTbFilter { float cutoff, resonance, drive, b0, z[4], k, r, g, in, out; KbOnePole feedback; };
struct KbTbVoice { KbOsm saw, square; KbEnv adsr; KbTbFilter filter; KbEnv env; float f; };
// examples/SynTHX.k:8-183
struct KbSxPartial { KbOsm osc; float f0, range, seed; int right; };
struct KbSxAdditive { KbSxPartial partial[4][3]; float frequency; };
struct KbSxVoice { KbSxAdditive notes[11]; KbEnv adsr; };

// examples/FM.k:27-74: three Operator<Fast::Sine> (klang.h:4140-4173) in series
struct KbFmOp { KbFastSine osc; KbEnv env; float amp, in; };
struct KbFmVoice { KbFmOp op[3]; KbEnv adsr; };

// examples/Subtractive/{Breakpoint,Ramp,Release}.k: a Fast::Sine times one breakpoint envelope
struct KbSenvVoice { KbFastSine osc; KbEnv env; int stop_when_finished; };

// examples/Additive/{Saw,Square}.k: 32 Fast::Sine partials
struct KbAddVoice { KbFastSine osc[32]; int square; /* 0 Saw.k: all partials, 1 Square.k: odd ones below Nyquist, 2 Nyquist.k: all below Nyquist */ };

// examples/Modulation/{AM,FM,FM2}.k: sine carrier, one or two sine modulators, ADSR
struct KbSmodVoice { KbFastSine carrier, mod1, mod2; KbEnv adsr; float f0; int graph; };

// ------------------------------------------------------------------ effect instances (graphs)
struct KbFxHdr { KbControl controls[KB_MAX_CONTROLS]; float cached[KB_MAX_CONTROLS]; };
// examples/PingPong.k
struct KbPingPong { KbDelay left, right; KbBasicOsc lfo; KbBiquad dc[2]; float delay; };
// examples/Reverb.k
#define KB_RV_MAXREFL 20
struct KbRvFDelay { KbDelay delay; KbBiquad filter; float gain, in, out; };
struct KbRvLate { KbRvFDelay d[4]; float in, out; };
struct KbReverb {
	KbDelay dl, dr;
	int count; float times[KB_RV_MAXREFL], gl[KB_RV_MAXREFL], gr[KB_RV_MAXREFL];
	float length, size;
	KbBiquad lpf[2], hpf[2];
	KbRvLate mid[2], late[2];
};
// examples/Delay/PingPong.k, examples/Delay/Reverb.k
struct KbDPingPong { KbDelay l, r; };
struct KbDReverb { KbDelay feedforward, feedback; KbBiquad filter; float out; };
struct KbGainFx { int unused; };
// examples/Gain/RM.k, Tremolo.k: the LFO (Pan.k and Clipping.k carry no state and use KbGainFx)
struct KbLfoFx { KbFastSine lfo; };
// examples/Delay/Echo.k, Feedback.k: one Delay<192000>
struct KbOneDelayFx { KbDelay delay; };
// examples/Filtering/IIR.k: the smoother's last output
struct KbIirFx { float last; };
// examples/Filtering/WahWah.k: Biquad::LPF + Fast::Sine LFO
struct KbWahWahFx { KbBiquad lpf; KbFastSine lfo; };
// examples/Modulation/{Flanger,ModDelay,Chorus}.k: one Delay<192000>, sine LFOs / a triangle LFO
struct KbModDelayFx { KbDelay delay; KbFastSine lfo[3]; KbOsm tri; };
