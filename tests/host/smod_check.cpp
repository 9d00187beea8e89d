// Host-side check (g++, no CUDA): the product's Modulation/AM.k, FM.k and FM2.k voice (kb_smod_on / kb_smod_tick of
// klang_b200/csrc/kb_graphs.cuh, the functions kb_voice_kernel<KB_SY_AM> runs per lane) rendered on the host: note on, a release, a
// second note on the same voice (AM.k keeps its modulator's phase, the FM programs reset their oscillators).  Raw float32 to stdout;
// tests/test_host_logic.py compares with the oracle bit for bit.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include "../../klang_b200/csrc/kb_graphs.cuh"

static void scenario(int graph, float fs_hz, float c0, float c1, float c2) {
	const KbFs fs = kb_make_fs(fs_hz);
	KbSmodVoice v;
	memset(&v, 0, sizeof(v));
	kb_smod_construct(fs, graph, v);
	int stage = KB_NOTE_OFF;
	const int marks[4] = { 0, 1500, 2100, 4000 };                           // on(60) | release | on(67) | end
	for (int seg = 0; seg < 3; seg++) {
		if (seg == 0) { kb_smod_on(fs, v, 60.f); stage = KB_NOTE_SUSTAIN; }
		if (seg == 1 && stage != KB_NOTE_OFF) { kb_adsr_release(fs, v.adsr); stage = KB_NOTE_RELEASE; }
		if (seg == 2) { kb_smod_on(fs, v, 67.f); stage = KB_NOTE_SUSTAIN; }
		const bool active = stage != KB_NOTE_OFF;                             // a voice active at the start of a block is ticked to its end
		for (int s = marks[seg]; s < marks[seg + 1]; s++) {
			const float y = active ? kb_smod_tick(fs, c0, c1, c2, v, stage) : 0.f;
			fwrite(&y, sizeof(float), 1, stdout);
		}
	}
}

int main() {
	scenario(KB_SY_AM, 48000.f, 2.2f, 0.9f, 0.f);
	scenario(KB_SY_MOD_FM, 44100.f, 1.5f, 7.0f, 0.f);
	scenario(KB_SY_MOD_FM2, 48000.f, 3.0f, 10.0f, 6.791f);
	return 0;
}
