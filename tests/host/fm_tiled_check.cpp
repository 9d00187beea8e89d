// Host-side check (g++, no CUDA): the time-parallel form of the FM.k voice — envelope rows from the run-length envelope
// (kb_envr_run) in tiles of 128 ticks, every sample from kb_fm_at(t), kb_fm_block_end at the end of the block — equals the
// per-tick kb_fm_tick bit for bit, samples AND the voice state left behind, over ragged blocks, releases that fall inside a
// tile, notes that run into Off and re-triggers.  These are the functions kb_fm_tiled_kernel runs (klang_b200/csrc/kb_tiled.cuh).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <vector>
#include "../../klang_b200/csrc/kb_graphs.cuh"

static const int TILE = 128;

// one block of `n` ticks through the tiled formulation
static void tiled_block(const KbFs& fs, float i1, float i2, KbFmVoice& v, int& stage, float* out, int n) {
	std::vector<float> row[4];
	KbEnv* envs[4] = { &v.op[0].env, &v.op[1].env, &v.op[2].env, &v.adsr };
	KbEnvR r[4];
	for (int e = 0; e < 4; e++) { row[e].resize(n); kb_envr_load(r[e], *envs[e]); }
	KbFmSample last;
	memset(&last, 0, sizeof(last));
	for (int base = 0; base < n; base += TILE) {
		const int steps = n - base < TILE ? n - base : TILE;
		for (int e = 0; e < 4; e++) kb_envr_run(fs, r[e], envs[e]->px, envs[e]->py, row[e].data() + base, steps);
		for (int t = steps - 1; t >= 0; t--) {                     // any order: samples are independent
			const KbFmSample s = kb_fm_at(v, (uint32_t)(base + t), i1, i2, row[0][base + t], row[1][base + t], row[2][base + t], row[3][base + t]);
			out[base + t] = s.out;
			if (base + t == n - 1) last = s;
		}
	}
	for (int e = 0; e < 4; e++) kb_envr_store(r[e], *envs[e]);
	kb_fm_block_end(v, (uint32_t)n, i1, i2, last);
	if (v.adsr.stage == KB_ENV_OFF) stage = KB_NOTE_OFF;
}

int main() {
	long long samples = 0, bad = 0, state_bad = 0, ended = 0;
	float peak = 0.f;
	unsigned seed = 12345u;
	auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return seed >> 8; };
	const float rates[3] = { 44100.f, 48000.f, 96000.f };
	for (int trial = 0; trial < 60; trial++) {
		const KbFs fs = kb_make_fs(rates[trial % 3]);
		KbControl c[4] = { kb_dial(0.001f, 10.f, 1.0f), kb_dial(0.f, 10.f, 0.37f), kb_dial(0.f, 10.f, 0.37f), kb_dial(0.f, 1.f, 0.5f) };
		if (trial) {
			kb_control_set(c[0], 0.001f + (rnd() % 1000) * 0.01f);
			kb_control_set(c[1], (rnd() % 1000) * 0.01f);
			kb_control_set(c[2], (rnd() % 1000) * 0.01f);
			kb_control_set(c[3], (rnd() % 1000) * (trial % 4 == 1 ? 0.00001f : 0.001f));
		}
		KbFmVoice a, b;
		kb_fm_construct(fs, a);
		kb_fm_on(fs, c, a, 30.f + (float)(rnd() % 70));
		memcpy(&b, &a, sizeof(a));
		int sa = KB_NOTE_SUSTAIN, sb = KB_NOTE_SUSTAIN;
		const int nblocks = 22 + (int)(rnd() % 8);
		const int release_block = 1 + (int)(rnd() % 6), retrigger_block = trial % 5 == 2 ? release_block + 2 : -1;
		for (int k = 0; k < nblocks; k++) {
			const int sizes[8] = { 1, 7, 117, 128, 129, 300, 1000, 4096 };
			const int n = (trial % 7 == 3 || (trial % 2 == 0 && k > release_block + 1)) ? 4096 : sizes[rnd() % 8];
			if (k == release_block) { kb_adsr_release(fs, a.adsr); kb_adsr_release(fs, b.adsr); }
			if (k == retrigger_block) {
				const float p = 40.f + (float)(rnd() % 40);
				kb_fm_on(fs, c, a, p); kb_fm_on(fs, c, b, p); sa = sb = KB_NOTE_SUSTAIN;
			}
			if (k % 5 == 4) { kb_control_set(c[1], (rnd() % 1000) * 0.01f); kb_control_set(c[2], (rnd() % 1000) * 0.01f); }
			std::vector<float> ya(n), yb(n);
			// the lane-per-voice kernel ticks a voice for the whole block once it was active at the block start (kb_voice_kernel)
			if (sa != KB_NOTE_OFF) { for (int t = 0; t < n; t++) ya[t] = kb_fm_tick(fs, c[1].value, c[2].value, a, sa); }
			if (sb != KB_NOTE_OFF) tiled_block(fs, c[1].value, c[2].value, b, sb, yb.data(), n);
			for (int t = 0; t < n; t++) { samples++; if (kb_fbits(ya[t]) != kb_fbits(yb[t])) bad++; if (fabsf(ya[t]) > peak) peak = fabsf(ya[t]); }
			if (memcmp(&a, &b, sizeof(a)) != 0 || sa != sb) state_bad++;
		}
		if (sa == KB_NOTE_OFF) ended++;
	}
	printf("fm tiled form: %lld samples, %lld mismatches, %lld state mismatches, %lld of 60 notes ran into Off, peak %g\n", samples, bad, state_bad, ended, peak);
	return (bad || state_bad || ended < 10 || ended > 55 || peak < 0.05f) ? 1 : 0;
}
