// Host-side check (g++, no CUDA): the product's Breakpoint.k / Ramp.k / Release.k voice (kb_senv_on / kb_senv_tick of
// klang_b200/csrc/kb_graphs.cuh, the functions kb_voice_kernel<KB_SY_BREAKPOINT> runs per lane) rendered on the host.  Prints the
// samples of one scenario per graph as raw float32; tests/test_host_logic.py compares them bit for bit with the oracle.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include "../../klang_b200/csrc/kb_graphs.cuh"

static void scenario(int graph, float fs_hz, float pitch, const KbControl* c, int n, int release_at) {
	const KbFs fs = kb_make_fs(fs_hz);
	KbSenvVoice v;
	kb_senv_construct(fs, graph, v);
	kb_senv_on(fs, graph, c, v, pitch);
	int stage = KB_NOTE_SUSTAIN;
	bool active = true;                                                   // a voice active at the start of a block is ticked to its end
	for (int s = 0; s < n; s++) {                                         // (Note::process(buffer), klang.h:4295-4303; kb_voice_kernel)
		if (s == release_at) {                                            // block boundary: NoteBase::release -> off()
			if (graph == KB_SY_RELEASE) kb_env_release(fs, v.env, c[3].value, 0.f); else stage = KB_NOTE_OFF;
			active = stage != KB_NOTE_OFF;
		}
		const float y = active ? kb_senv_tick(fs, v, stage) : 0.f;
		fwrite(&y, sizeof(float), 1, stdout);
	}
}

int main() {
	KbControl bp[2] = { kb_dial(0.05f, 1.0f, 0.05f), kb_dial(0.1f, 1.0f, 0.1f) };
	KbControl rp[1] = { kb_dial(0.1f, 1.0f, 0.1f) };
	KbControl rl[4] = { kb_dial(0.f, 1.f, 0.002f), kb_dial(0.f, 1.f, 0.1f), kb_dial(0.f, 1.f, 0.05f), kb_dial(0.f, 1.f, 1.0f) };
	kb_control_set(rl[1], 0.01f); kb_control_set(rl[2], 0.5f); kb_control_set(rl[3], 0.02f);
	scenario(KB_SY_BREAKPOINT, 48000.f, 60.f, bp, 6000, 5500);
	scenario(KB_SY_RAMP, 44100.f, 72.f, rp, 6000, -1);
	scenario(KB_SY_RELEASE, 48000.f, 45.f, rl, 4000, 1500);
	return 0;
}
