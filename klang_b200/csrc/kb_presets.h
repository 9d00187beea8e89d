// klang-b200 — factory presets of the bound programs (Plugin::presets, klang.h:1940-1981, 4195-4200): name and control values as each
// program's constructor lists them.  A host loads a preset by writing its values through Control::set like any parameter change
// (klang.h:1725-1728, 4444-4447) and then calls Controller::onPreset (klang.h:4190, 4415-4420), whose preset() hook none of the bound
// programs overrides.  tests/test_host_logic.py checks this table against the compiled reference.
#pragma once
#include "kb_state.h"

struct KbPreset { int is_synth, graph; const char* name; int count; float values[KB_MAX_CONTROLS]; };
static const KbPreset kb_presets[] = {
	// examples/PingPong.k:23-32
	{ 0, KB_FX_PINGPONG, "Phat + Sinister", 6, { 0.958f, 0.018f, 0.166f, 0.579f, 0.001f, 0.026f } },
	{ 0, KB_FX_PINGPONG, "Funky Beat", 6, { 0.663f, 0.248f, 0.411f, 0.594f, 2.000f, 0.283f } },
	{ 0, KB_FX_PINGPONG, "Station Concourse", 6, { 0.584f, 0.380f, 0.000f, 0.010f, 0.775f, 0.380f } },
	{ 0, KB_FX_PINGPONG, "Metal Voice", 6, { 0.940f, 0.025f, 0.000f, 0.010f, 0.138f, 0.025f } },
	{ 0, KB_FX_PINGPONG, "Bad Trip", 6, { 0.881f, 0.651f, 0.560f, 0.028f, 0.138f, 0.772f } },
	{ 0, KB_FX_PINGPONG, "Pitchy + Scratchy", 6, { 0.881f, 0.643f, 0.4f, 0.127f, 0.001f, 0.500f } },
	{ 0, KB_FX_PINGPONG, "Burpy Bubbles", 6, { 0.272f, 0.234f, 0.648f, 0.127f, 0.001f, 0.201f } },
	{ 0, KB_FX_PINGPONG, "Doctor Who?", 6, { 0.325f, 0.008f, 0.382f, 1.000f, 0.001f, 0.000f } },
	// examples/Reverb.k:113-115
	{ 0, KB_FX_REVERB, "Large Hall", 10, { 1.000f, 0.000f, 0.419f, 0.329f, 1.000f, 10.000f, 100.000f, 0.500f, 0.500f, 0.100f } },
	// examples/SuperSaw.k:45-50
	{ 1, KB_SY_SUPERSAW, "Pluck", 3, { 0.001f, 0.615f, 0.098f } },
	{ 1, KB_SY_SUPERSAW, "Trance Lead", 3, { 0.001f, 0.1f, 0.6f } },
	{ 1, KB_SY_SUPERSAW, "Synth Pad", 3, { 1.000f, 0.037f, 0.167f } },
	{ 1, KB_SY_SUPERSAW, "Paris Traffic", 3, { 0.126f, 0.100f, 1.000f } },
	// examples/Modulation/FM2.k:47-52
	{ 1, KB_SY_MOD_FM2, "Violin", 3, { 3.000f, 10.000f, 6.791f } },
	{ 1, KB_SY_MOD_FM2, "Cello", 3, { 1.490f, 7.076f, 1.523f } },
	{ 1, KB_SY_MOD_FM2, "Oboe", 3, { 3.000f, 0.755f, 10.000f } },
	{ 1, KB_SY_MOD_FM2, "Harmonica", 3, { 2.500f, 4.900f, 8.443f } },
};
static inline const KbPreset* kb_preset_find(int is_synth, int graph, int index) {
	for (const KbPreset& p : kb_presets) if (p.is_synth == is_synth && p.graph == graph && index-- == 0) return &p;
	return nullptr;
}
static inline int kb_preset_count(int is_synth, int graph) {
	int n = 0;
	for (const KbPreset& p : kb_presets) n += p.is_synth == is_synth && p.graph == graph;
	return n;
}
