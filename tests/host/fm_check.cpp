// Host-side check (g++, no CUDA): the product's FM.k voice (kb_fm_on / kb_fm_tick of klang_b200/csrc/kb_graphs.cuh, the functions
// the device kernel runs) rendered on the host.  Prints the samples of two scenarios as raw float32; tests/test_host_logic.py
// compares them bit for bit with the oracle.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include "../../klang_b200/csrc/kb_graphs.cuh"

static void scenario(float fs_hz, float pitch, const float* ctl, int n, int release_at) {
	const KbFs fs = kb_make_fs(fs_hz);
	KbControl c[4] = { kb_dial(0.001f, 10.f, 1.0f), kb_dial(0.f, 10.f, 0.37f), kb_dial(0.f, 10.f, 0.37f), kb_dial(0.f, 1.f, 0.5f) };
	for (int k = 0; k < 4; k++) kb_control_set(c[k], ctl[k]);
	KbFmVoice v;
	kb_fm_construct(fs, v);
	kb_fm_on(fs, c, v, pitch);
	int stage = KB_NOTE_SUSTAIN;
	for (int s = 0; s < n; s++) {
		if (s == release_at) kb_adsr_release(fs, v.adsr);
		const float y = (stage == KB_NOTE_OFF) ? 0.f : kb_fm_tick(fs, c[1].value, c[2].value, v, stage);
		fwrite(&y, sizeof(float), 1, stdout);
	}
}

int main() {
	const float a[4] = { 1.0f, 0.37f, 0.37f, 0.5f }, b[4] = { 2.5f, 3.0f, 7.5f, 0.002f };
	scenario(48000.f, 60.f, a, 3000, 1500);
	scenario(44100.f, 72.f, b, 2000, 700);
	return 0;
}
