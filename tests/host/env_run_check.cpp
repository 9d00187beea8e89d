// Host-side check (g++, no CUDA): the run-length envelope kb_envr_run and the closed-form oscillator kb_osm_at /
// kb_osm_advance of klang_b200/csrc/kb_prims.cuh are bit-identical to the per-tick forms, over random envelopes
// (breakpoints, loops, releases at random times, ragged chunking) and random oscillator settings.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include "../../klang_b200/csrc/kb_prims.cuh"

static uint64_t rng_state = 88172645463325252ull;
static uint32_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t)(rng_state >> 11); }
static float frand(float lo, float hi) { return lo + (hi - lo) * (float)(rnd() & 0xffffff) / 16777216.f; }
static uint32_t bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

int main() {
	long long checked = 0, bad = 0;
	for (int trial = 0; trial < 6000; trial++) {
		const KbFs fs = kb_make_fs(trial & 1 ? 48000.f : 44100.f);
		KbEnv a;
		if (trial % 3 == 0) { kb_adsr_construct(fs, a); kb_adsr_set(fs, a, frand(0, 0.02f), frand(0, 0.02f), frand(0, 1), frand(0, 0.02f)); }
		else {
			kb_env_construct(fs, a);
			int np = 1 + rnd() % 6; float xy[32]; float x = 0;
			for (int p = 0; p < np; p++) { xy[2 * p] = x; xy[2 * p + 1] = (rnd() % 8 == 0) ? ((rnd() & 1) ? -0.f : 0.f) : frand(-2, 2000); x += (rnd() % 5 == 0) ? 0.f : frand(0.0001f, (rnd() % 4 == 0) ? 0.05f : 0.01f); }
			kb_env_set_points(fs, a, np, xy);
			if (rnd() % 3 == 0) { int s = rnd() % np, t = s + rnd() % (np - s); kb_env_set_loop(a, s, t); }
		}
		KbEnv b = a; const KbEnv b0 = a;
		const int n = 200 + rnd() % 3000, release_at = (rnd() % 2) ? (int)(rnd() % n) : -1;
		const float rel_time = frand(0, 0.01f), rel_level = frand(0, 1);
		static float ref[4096], got[4096];
		for (int s = 0; s < n; s++) {
			if (s == release_at) { if (trial % 3 == 0) kb_adsr_release(fs, a); else kb_env_release(fs, a, rel_time, rel_level); }
			ref[s] = kb_env_tick(fs, a);
		}
		for (int s = 0; s < n;) {
			int len = 1 + rnd() % 200;
			if (len > n - s) len = n - s;
			if (release_at > s && release_at < s + len) len = release_at - s;       // events fall on chunk boundaries
			if (s == release_at) { if (trial % 3 == 0) kb_adsr_release(fs, b); else kb_env_release(fs, b, rel_time, rel_level); }
			KbEnvR r; kb_envr_load(r, b);
			kb_envr_run(fs, r, b.px, b.py, got + s, len);
			kb_envr_store(r, b);
			s += len;
		}
		for (int form = 0; form < 2; form++) {	// the uniform-group form (kb_envr_run16) and the whole-tile form (kb_envr_run_tile), in tiles of at most 128 ticks
			KbEnv c = b0; static float got16[4096];
			for (int s = 0; s < n;) {
				int len = (trial & 4) ? 128 : (trial & 8) ? 16 * (1 + rnd() % 8) : 1 + rnd() % 200;
				if (len > n - s) len = n - s;
				if (release_at > s && release_at < s + len) len = release_at - s;
				if (s == release_at) { if (trial % 3 == 0) kb_adsr_release(fs, c); else kb_env_release(fs, c, rel_time, rel_level); }
				KbEnvR r; kb_envr_load(r, c);
				if (form) kb_envr_run_tile<false>(fs, r, c.px, c.py, got16 + s, len); else kb_envr_run16<false>(fs, r, c.px, c.py, got16 + s, len);
				kb_envr_store(r, c);
				s += len;
			}
			for (int i = 0; i < n; i++) { checked++; if (bits(ref[i]) != bits(got16[i])) { if (bad < 5) printf("env16 trial %d i %d: %a vs %a\n", trial, i, ref[i], got16[i]); bad++; } }
			if (bits(a.time) != bits(c.time) || a.stage != c.stage || a.point != c.point || bits(a.r_out) != bits(c.r_out) || a.r_active != c.r_active || bits(a.out) != bits(c.out)) {
				if (bad < 5) printf("env16 trial %d final state differs\n", trial);
				bad++;
			}
		}
		for (int i = 0; i < n; i++) { checked++; if (bits(ref[i]) != bits(got[i])) { if (bad < 5) printf("env trial %d i %d: %a vs %a\n", trial, i, ref[i], got[i]); bad++; } }
		if (bits(a.time) != bits(b.time) || a.stage != b.stage || a.point != b.point || bits(a.r_out) != bits(b.r_out) || a.r_active != b.r_active || bits(a.out) != bits(b.out)) {
			if (bad < 5) printf("env trial %d final state differs\n", trial);
			bad++;
		}
	}
	// oscillators: closed form vs sequential
	for (int trial = 0; trial < 2000; trial++) {
		const KbFs fs = kb_make_fs(48000.f);
		KbOsm o; kb_osm_construct(o, trial & 1, (trial & 1) ? frand(0, 1) : ((trial & 2) ? 0.f : frand(0, 1)));
		kb_osm_set_fpd(fs, o, frand(20, 20000), frand(0, 6.28f), (trial & 4) ? frand(0, 1) : 0.f);
		KbOsm seq = o;
		const int n = 1 + rnd() % 700;
		for (int i = 0; i < n; i++) { const float y = kb_osm_tick(seq), z = kb_osm_at(o, (uint32_t)i); checked++; if (bits(y) != bits(z)) { if (bad < 10) printf("osm trial %d i %d: %a vs %a\n", trial, i, y, z); bad++; } }
		KbOsm adv = o; kb_osm_advance(adv, (uint32_t)n);
		if (adv.offset != seq.offset || adv.state != seq.state) { if (bad < 10) printf("osm trial %d advance differs\n", trial); bad++; }
	}
	printf("checked %lld values, %lld mismatches\n", checked, bad);
	return bad != 0;
}
