// klang-b200 — C++ convenience layer over the C ABI (include/klang_b200.h), shaped like the reference's block drivers.
//
// klang's host-facing surface is a pair of C++ classes (klang.h:4203-4217 Effect, 4376-4467 Synth, 4703-4859 Stereo::*):
// a host owns an object, pushes control values and MIDI between blocks, and calls process() with caller-owned planar
// float buffers.  The classes below keep exactly that surface — names, argument meaning, in-place / overwrite
// semantics, "no exceptions from process()" — but every block is evaluated by libklang_b200.so on a B200.
// They are the shim a klang.h-compatible header forwards to; header-only, C++11.
#pragma once
#include <algorithm>
#include <stdexcept>
#include <string>
#include <vector>

#include "klang_b200.h"

namespace klang_b200 {

struct Error : std::runtime_error { explicit Error(const std::string& what) : std::runtime_error(what + ": " + kb_last_error()) {} };

// N independent instances of one Effect / Stereo::Effect graph (N = 1 is the reference's single plugin object).
class Effect {
public:
	Effect(int graph, float sample_rate, int max_block, int instances = 1, int device = 0)
		: bank_(kb_fx_bank_create(graph, instances, sample_rate, max_block, device)) { if (!bank_) throw Error("kb_fx_bank_create"); }
	~Effect() { kb_fx_bank_destroy(bank_); }
	Effect(const Effect&) = delete;
	Effect& operator=(const Effect&) = delete;

	int channels() const { return kb_fx_bank_channels(bank_); }
	int instances() const { return kb_fx_bank_instances(bank_); }
	// controls[index].set(value)                                              klang.h:1725-1728
	void setControl(int index, float value, int instance = 0) { kb_fx_bank_set_control(bank_, instance, index, value); }
	float control(int index, int instance = 0) { float v = 0.f; kb_fx_bank_get_control(bank_, instance, index, &v); return v; }
	// Effect::process(buffer): in place, planar [instances][channels][length]  klang.h:4208-4216, 4708-4716
	bool process(float* buffer, int length) noexcept { return kb_fx_bank_process(bank_, buffer, length, 0) == KB_OK; }
	// mono plugin object: one channel pointer; stereo plugin object: two planar channel pointers (JUCE style)
	bool process(float* const* channelData, int length) {
		const int C = channels();
		if (C == 1 || channelData[1] == channelData[0] + length) return process(channelData[0], length);
		scratch_.resize((size_t)C * length);
		for (int c = 0; c < C; c++) std::copy(channelData[c], channelData[c] + length, scratch_.begin() + (size_t)c * length);
		const bool ok = process(scratch_.data(), length);
		for (int c = 0; c < C; c++) std::copy(scratch_.begin() + (size_t)c * length, scratch_.begin() + (size_t)(c + 1) * length, channelData[c]);
		return ok;
	}
	kb_fx_bank* handle() { return bank_; }
private:
	kb_fx_bank* bank_;
	std::vector<float> scratch_;
};

// N independent Synth objects of one graph with `voices` notes each.
class Synth {
public:
	Synth(int graph, float sample_rate, int max_block, int voices = 32, int instances = 1, int device = 0)
		: bank_(kb_synth_bank_create(graph, instances, voices, sample_rate, max_block, device)) { if (!bank_) throw Error("kb_synth_bank_create"); }
	~Synth() { kb_synth_bank_destroy(bank_); }
	Synth(const Synth&) = delete;
	Synth& operator=(const Synth&) = delete;

	int channels() const { return kb_synth_bank_channels(bank_); }
	int voices() const { return kb_synth_bank_voices(bank_); }
	void setControl(int index, float value, int instance = 0) { kb_synth_bank_set_control(bank_, instance, index, value); }
	// Synth::noteOn / noteOff                                                  klang.h:4423-4434
	int noteOn(int pitch, float velocity, int instance = 0) { return kb_synth_bank_note_on(bank_, instance, pitch, velocity); }
	void noteOff(int pitch, float velocity = 0.f, int instance = 0) { kb_synth_bank_note_off(bank_, instance, pitch, velocity); }
	// Synth::input(status, byte1, byte2): raw MIDI                             templates/juce/synth/Source/klang.h:3921-3929
	void input(int status, int byte1, int byte2, int instance = 0) { kb_synth_bank_midi(bank_, instance, status, byte1, byte2); }
	// Synth::process(float* buffer, int length) / Stereo::Synth::process: the output block is overwritten,
	// planar [instances][channels][length]                                     klang.h:4440-4466, 4830-4858
	bool process(float* buffer, int length, unsigned flags = 0) noexcept { return kb_synth_bank_process(bank_, buffer, length, flags) == KB_OK; }
	kb_synth_bank* handle() { return bank_; }
private:
	kb_synth_bank* bank_;
};

}  // namespace klang_b200
