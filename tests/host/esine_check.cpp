// Host-side check (g++, no CUDA): the time-parallel form of the one-envelope voices — Breakpoint.k / Ramp.k / Release.k (KbSenvVoice) and
// Modulation/AM.k (KbSmodVoice) — as kb_esine_tiled_kernel runs it: kb_es_begin at the block start, envelope rows from the run-length
// envelope (kb_envr_run) in tiles of 128 ticks, every sample from kb_es_at(t) in any order, kb_es_end + the envelope store at the end.
// It must equal the per-tick forms (kb_senv_tick / kb_smod_tick) bit for bit, samples AND the state left behind, over ragged blocks,
// releases, notes running into Off, re-triggers and control changes between blocks.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <vector>
#include "../../klang_b200/csrc/kb_graphs.cuh"

static const int TILE = 128;
static unsigned seed = 777u;
static unsigned rnd() { seed = seed * 1664525u + 1013904223u; return seed >> 8; }

template <class VOICE>
static void tiled_block(const KbFs& fs, float c0, float c1, VOICE& v, int& stage, float* out, int n) {
	kb_es_begin(fs, v, c0);
	KbEnv& e = kb_es_env(v);
	KbEnvR r; kb_envr_load(r, e);
	std::vector<float> row(n);
	for (int base = 0; base < n; base += TILE) {
		const int steps = n - base < TILE ? n - base : TILE;
		kb_envr_run(fs, r, e.px, e.py, row.data() + base, steps);
		for (int t = steps - 1; t >= 0; t--) out[base + t] = kb_es_at(v, (uint32_t)(base + t), c1, row[base + t]);
	}
	kb_envr_store(r, e);
	VOICE m = v;
	kb_es_end(m, (uint32_t)n);
	kb_es_writeback(v, m);
	if (kb_es_stops(v) && e.stage == KB_ENV_OFF) stage = KB_NOTE_OFF;
}

int main() {
	long long samples = 0, bad = 0, state_bad = 0, ended = 0;
	float peak = 0.f;
	const float rates[3] = { 44100.f, 48000.f, 96000.f };
	const int sizes[8] = { 1, 7, 117, 128, 129, 300, 1000, 4096 };
	for (int trial = 0; trial < 80; trial++) {
		const KbFs fs = kb_make_fs(rates[trial % 3]);
		const int kind = trial % 4;                                            // 0 Breakpoint, 1 Ramp, 2 Release, 3 AM
		const int graph = kind == 0 ? KB_SY_BREAKPOINT : kind == 1 ? KB_SY_RAMP : kind == 2 ? KB_SY_RELEASE : KB_SY_AM;
		KbControl c[4] = { kb_dial(0.f, 1.f, 0.002f + (rnd() % 100) * 0.001f), kb_dial(0.f, 1.f, (rnd() % 100) * 0.002f), kb_dial(0.f, 1.f, (rnd() % 100) * 0.01f),
		                   kb_dial(0.f, 1.f, 0.01f + (rnd() % 100) * 0.001f) };
		float am0 = 0.01f + (rnd() % 300) * 0.01f, am1 = (rnd() % 100) * 0.01f;
		KbSenvVoice sa, sb; KbSmodVoice ma, mb;
		memset(&sa, 0, sizeof(sa)); memset(&ma, 0, sizeof(ma));
		if (kind < 3) { kb_senv_construct(fs, graph, sa); kb_senv_on(fs, graph, c, sa, 30.f + (float)(rnd() % 70)); sb = sa; }
		else { kb_smod_construct(fs, graph, ma); kb_smod_on(fs, ma, 30.f + (float)(rnd() % 70)); mb = ma; }
		int st_a = KB_NOTE_SUSTAIN, st_b = KB_NOTE_SUSTAIN;
		const int nblocks = 10 + (int)(rnd() % 8), release_block = 1 + (int)(rnd() % 5), retrigger = trial % 5 == 2 ? release_block + 2 : -1;
		for (int k = 0; k < nblocks; k++) {
			const int n = (trial % 2 == 0 && k > release_block) ? 4096 : sizes[rnd() % 8];
			if (k == release_block) {
				if (kind == 2) { kb_env_release(fs, sa.env, c[3].value, 0.f); kb_env_release(fs, sb.env, c[3].value, 0.f); }
				else if (kind == 3) { kb_adsr_release(fs, ma.adsr); kb_adsr_release(fs, mb.adsr); }
			}
			if (k == retrigger) {
				const float p = 40.f + (float)(rnd() % 40);
				if (kind < 3) { kb_senv_on(fs, graph, c, sa, p); kb_senv_on(fs, graph, c, sb, p); } else { kb_smod_on(fs, ma, p); kb_smod_on(fs, mb, p); }
				st_a = st_b = KB_NOTE_SUSTAIN;
			}
			if (k % 4 == 3) { am0 = 0.01f + (rnd() % 300) * 0.01f; am1 = (rnd() % 100) * 0.01f; }
			std::vector<float> ya(n, 0.f), yb(n, 0.f);
			if (st_a != KB_NOTE_OFF) for (int t = 0; t < n; t++) ya[t] = kind < 3 ? kb_senv_tick(fs, sa, st_a) : kb_smod_tick(fs, am0, am1, 0.f, ma, st_a);
			if (st_b != KB_NOTE_OFF) { if (kind < 3) tiled_block(fs, 0.f, 0.f, sb, st_b, yb.data(), n); else tiled_block(fs, am0, am1, mb, st_b, yb.data(), n); }
			for (int t = 0; t < n; t++) { samples++; if (kb_fbits(ya[t]) != kb_fbits(yb[t])) bad++; if (fabsf(ya[t]) > peak) peak = fabsf(ya[t]); }
			if ((kind < 3 ? memcmp(&sa, &sb, sizeof(sa)) : memcmp(&ma, &mb, sizeof(ma))) != 0 || st_a != st_b) state_bad++;
		}
		if (st_a == KB_NOTE_OFF) ended++;
	}
	printf("one-envelope time-parallel form: %lld samples, %lld mismatches, %lld state mismatches, %lld of 80 notes ran into Off, peak %g\n", samples, bad, state_bad, ended, peak);
	return (bad || state_bad || ended < 10 || peak < 0.3f) ? 1 : 0;
}
